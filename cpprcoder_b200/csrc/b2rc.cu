// b2rc.cu -- the extern "C" layer of libb2rc.so (declared in include/b2rc.h).
//
// Host side of the drop-in boundary: argument checking, device scratch, kernel
// launches, the B2RC container.  No torch types, no CPU coding path: every call
// that codes bytes launches the sm_100a kernels in b2rc_kernels.cuh.
#include "../../include/b2rc.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <chrono>
#include <thread>
#include <vector>

#include "b2rc_kernels.cuh"
#include "b2rc_encseg.cuh"
#include "b2rc_adaptseg.cuh"
#include "b2rc_ans.cuh"
#include "b2rc_blk.cuh"

using namespace b2rc;

#define B2RC_PIPE_STREAMS 16
#define B2RC_PIPE_CHUNKS 16
#define B2RC_PIPE_MIN_CHUNK (64ull << 20)  // bytes of input per chunk, at least
#define B2RC_PHASES 8                       // launches per chunk of the phased static decode, at most
#define B2RC_PHASE_MIN_SYMS 16384u          // symbols per block and launch, at least

#define B2RC_MAX_DEVICES 16
struct b2rc_ctx {
    int device;
    int ndev;                          // > 1: the host-pointer calls shard over sub[0 .. ndev)
    int in_multi;                      // set while a sharded call runs: sub[0] is this context, its share is a plain call
    b2rc_ctx* sub[B2RC_MAX_DEVICES];   // sub[0] is this context itself
    cudaStream_t stream;
    // device scratch, grown on demand
    u8* slots;
    size_t slots_cap;
    u32* sizes;
    size_t sizes_cap;
    u16* freq16;
    size_t freq_cap;
    u8* stage_in;
    size_t stage_in_cap;
    u8* stage_out;
    size_t stage_out_cap;
    int* d_err;
    u64* d_total;
    // chunked host pipeline: a few streams, per-chunk "scan done" / "chunk done" events,
    // a device array of running payload ends (one per chunk) mirrored in pinned host memory
    cudaStream_t pipe[B2RC_PIPE_STREAMS];
    cudaEvent_t scan_done[B2RC_PIPE_CHUNKS], chunk_done[B2RC_PIPE_CHUNKS], index_ready;
    // phased static decode (b2rc_decode): per (chunk, phase) "kernel done" events, one D2H stream
    // per phase index, the parked coder state of every block
    cudaEvent_t phase_done[B2RC_PIPE_CHUNKS][B2RC_PHASES];
    cudaStream_t d2h[B2RC_PHASES];
    u32* dec_state;
    size_t dec_state_cap;
    u8* dec_model;  // adaptive coder: the count tables of every warp of blocks between launches
    size_t dec_model_cap;
    u64* d_ends;  // B2RC_PIPE_CHUNKS + 1
    u64* h_ends;  // pinned
    u64 max_chunks;  // <= B2RC_PIPE_CHUNKS; env B2RC_PIPE_CHUNKS overrides (tuning)
    u64 min_chunk;   // bytes of input per chunk, at least (B2RC_PIPE_MIN_CHUNK; env B2RC_PIPE_MIN_CHUNK: tests cut small streams)
    u32 ramp_chunks; // shorter chunks at both ends of the host pipeline (plan_chunks); env B2RC_PIPE_RAMP=0: equal chunks
    u64 max_phases;  // <= B2RC_PHASES; env B2RC_PHASES overrides (1 switches the phased decode off)
    u32 seg_syms;    // restart points of the byte rANS coder every so many symbols; env B2RC_RESTART_SYMS (0: none)
    u32 seg_syms_static;  // ... of the static range coder (the same env sets both)
    u32 dec_seg_warps;  // env B2RC_DEC_SEG_WARPS: warps (= segments) per CTA of k_dec_static_seg, 1 .. 12 (tuning)
    u32 ranges_warps;  // env B2RC_RANGES_WARPS: 1 = k_enc_ranges for 64 KiB blocks too, 2 / 3 = k_enc_ranges2, unset = by size
    u32 seg_auto;    // no env setting: streams of few blocks get more points per block (seg_for)
    u32 seg_force;   // b2rc_ctx_force_restart: this spacing whatever the stream (0: not forced)
    u32 seg_syms_adaptive;  // the same for the adaptive coder (524 B per point); env B2RC_ADAPTIVE_RESTART_SYMS (0: none)
    u32 seg_syms_adaptive_wide;  // ... for blocks above 65536 bytes (1036 B per point); env B2RC_ADAPTIVE_RESTART_SYMS_WIDE
    u32* restart;    // device scratch: the table while a container is being written / read
    size_t restart_cap;
    // segmented static encode (b2rc_encseg.cuh): the range pass's records and the segments' final lows
    u8* h_stage;  // pinned host staging of the *_staged calls
    size_t h_stage_cap;
    u8* h_in;     // pinned host staging of large pageable sources
    size_t h_in_cap;
    u32* seg_recs;
    size_t seg_recs_cap;
    u32* seg_lows;
    size_t seg_lows_cap;
    u32 enc_seg_syms;  // symbols per segment of the static encoder; env B2RC_ENC_SEG_SYMS (0: one chain per block, k_enc_static)
    u32 force_exact;   // env B2RC_FORCE_EXACT=1 (tests): every static block takes the reference-shaped path
    u32 adaptive_two_warps;  // env B2RC_ADAPTIVE_TWO_WARPS=1: k_enc_adaptive2 (model and coder on a warp each; slower, b2rc_adaptseg.cuh)
    struct Result {
        int err;
        int pad;
        u64 total;
        u8 header[B2RC_HEADER_BYTES];
    } * h_res;  // pinned
    u32* blk_rounds;  // block sort: doubling rounds per block of the last forward call
    size_t blk_rounds_cap;
    u32* blk_tie_list;  // block sort: [0] how many blocks of the call have a period, [1 ..] which
    size_t blk_tie_cap;
    u8* blk_tmp;        // block sort in front of a coder (b2rc_blkrc_*): the transformed stream between the two
    size_t blk_tmp_cap;
    u8* blk_ties;     // block sort: scratch of the tie replay (block list, ranks, range queues)
    size_t blk_ties_cap;
    u64 blk_last_blocks;
    u64 launches;
    char last_err[256];
    // optional per-kernel timing (b2rc_profile): CUDA events on the launching stream
    int profiling;
    cudaEvent_t ev[B2RC_K_COUNT][2];
    int ev_used[B2RC_K_COUNT];
};

namespace
{
bool cuda_ok(b2rc_ctx* c, cudaError_t e, const char* what)
{
    if(e == cudaSuccess) {
        return true;
    }
    if(c) {
        snprintf(c->last_err, sizeof c->last_err, "%s: %s", what, cudaGetErrorString(e));
    }
    cudaGetLastError();
    return false;
}
#define CK(call)                          \
    do {                                  \
        if(!cuda_ok(ctx, (call), #call)) { \
            return B2RC_E_CUDA;           \
        }                                 \
    } while(0)

template <class T>
int grow(b2rc_ctx* ctx, T*& p, size_t& cap, size_t want_bytes)
{
    if(want_bytes <= cap) {
        return B2RC_OK;
    }
    if(p) {
        CK(cudaFree(p));
        p = nullptr;
        cap = 0;
    }
    size_t bytes = want_bytes + want_bytes / 8 + 4096;
    void* q = nullptr;
    if(!cuda_ok(ctx, cudaMalloc(&q, bytes), "cudaMalloc")) {
        bytes = want_bytes;
        if(!cuda_ok(ctx, cudaMalloc(&q, bytes), "cudaMalloc")) {
            return B2RC_E_NOMEM;
        }
    }
    p = static_cast<T*>(q);
    cap = bytes;
    return B2RC_OK;
}

}  // namespace
extern "C" {
static int grow_host(b2rc_ctx* ctx, u8*& buf, size_t& cap, size_t want);  // pinned host staging, below
}
namespace
{
bool block_ok(u32 block)
{
    return block >= B2RC_MIN_BLOCK && block <= B2RC_MAX_BLOCK && (block % 64u) == 0;
}
bool mode_ok(int mode)
{
    return mode == B2RC_MODE_STATIC || mode == B2RC_MODE_ADAPTIVE || mode == B2RC_MODE_RANS_BYTE ||
           mode == B2RC_MODE_RANS_WORD;
}
bool is_ans(int mode)
{
    return mode == B2RC_MODE_RANS_BYTE || mode == B2RC_MODE_RANS_WORD;
}
bool seg_ok(u32 block_size, u32 seg_syms)
{
    return seg_syms >= 64u && (seg_syms % 64u) == 0u && seg_syms < block_size;
}
// segment length a container of this mode / block size is written with by this context
bool has_restart(int mode)
{
    // the one-chain-per-block coders: static model (a point is three words) and, with the model's
    // counts in every point, the adaptive range coder
    return mode == B2RC_MODE_STATIC || mode == B2RC_MODE_RANS_BYTE || mode == B2RC_MODE_ADAPTIVE;
}
// u32 words per restart point: {bytes shifted, low, range}; the adaptive coder adds its 256 symbol counts
// (u16, two per word; u32 for blocks above 65536 bytes)
u32 rec_words(int mode, u32 block_size)
{
    return mode != B2RC_MODE_ADAPTIVE ? 3u : (block_size <= 65536u ? ADAPT_REC_WORDS : ADAPT_REC_WORDS_WIDE);
}
// The decoder runs one warp per 32 blocks and segment; it wants about as many warps as the 1 GiB / 64 KiB case
// gives it at the default spacing (512 x 8, seven CTAs of four on every SM).  A stream with fewer blocks gets its
// points closer together, down to B2RC_MIN_RESTART_SYMS: 128 MiB decode in 0.64 instead of 1.42 ms, 16 MiB in
// 0.27 instead of 1.39 ms (tools/dec_perf.py) for at most +1.3 % size.  12 bytes per point.
constexpr u64 SEG_AUTO_WARPS = 4096;
u32 seg_for(const b2rc_ctx* ctx, int mode, u32 block_size, u64 nblocks)
{
    if(mode == B2RC_MODE_ADAPTIVE) {  // the points carry the model: one spacing per tree width, no halving
        const u32 seg = block_size <= 65536u ? ctx->seg_syms_adaptive : ctx->seg_syms_adaptive_wide;
        return seg_ok(block_size, seg) ? seg : 0u;
    }
    if(!has_restart(mode)) {
        return 0u;
    }
    if(ctx->seg_force) {
        return seg_ok(block_size, ctx->seg_force) ? ctx->seg_force : 0u;
    }
    u32 seg = mode == B2RC_MODE_STATIC ? ctx->seg_syms_static : ctx->seg_syms;
    if(!seg_ok(block_size, seg)) {
        return 0u;
    }
    if(ctx->seg_auto) {
        const u64 groups = (nblocks + 31u) / 32u;
        while(seg / 2u >= B2RC_MIN_RESTART_SYMS && (seg / 2u) % 64u == 0u &&
              groups * ((block_size + seg - 1u) / seg) < SEG_AUTO_WARPS) {
            seg /= 2u;
        }
    }
    return seg;
}
bool aligned16(const void* p)
{
    return ((uintptr_t)p & 15u) == 0;
}
u64 index_bytes(u64 nblocks)
{
    return (u64)B2RC_HEADER_BYTES + 8ull * (nblocks + 1);
}
// header flags: bit 0 = a restart table follows the payloads, bits 8.. = its segment length / 64
u32 flags_of(u32 seg_syms)
{
    return seg_syms ? (1u | ((seg_syms / 64u) << 8)) : 0u;
}
u64 align4(u64 x)
{
    return (x + 3ull) & ~3ull;
}

// dynamic shared memory of the coder kernels (one warp per CTA)
constexpr u32 smem_enc_static(bool wide)
{
    return (wide ? 257u : 256u) * 128u + 2u * TILE_BYTES;
}
constexpr u32 smem_enc_adaptive(bool wide)
{
    return 512u * 32u * (wide ? 4u : 2u) + 2u * TILE_BYTES;
}
constexpr u32 smem_dec_static()
{
    return DEC_STATIC_TAB + TILE_BYTES + INQ_BYTES;
}
constexpr u32 smem_dec_adaptive(bool wide)
{
    return 512u * 32u * (wide ? 4u : 2u) + TILE_BYTES + INQ_BYTES;
}

int set_smem_limits(b2rc_ctx* ctx)
{
    CK(cudaFuncSetAttribute(k_enc_static<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_enc_static(false)));
    CK(cudaFuncSetAttribute(k_enc_static<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_enc_static(true)));
    CK(cudaFuncSetAttribute(k_enc_adaptive<u16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_enc_adaptive(false)));
    CK(cudaFuncSetAttribute(k_enc_adaptive<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_enc_adaptive(true)));
    CK(cudaFuncSetAttribute(k_dec_static<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dec_static()));
    CK(cudaFuncSetAttribute(k_dec_static<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dec_static()));
    CK(cudaFuncSetAttribute(k_dec_adaptive<u16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dec_adaptive(false)));
    CK(cudaFuncSetAttribute(k_dec_adaptive<u16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dec_adaptive(false)));
    CK(cudaFuncSetAttribute(k_dec_adaptive<u32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dec_adaptive(true)));
    CK(cudaFuncSetAttribute(k_dec_adaptive<u32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dec_adaptive(true)));
    CK(cudaFuncSetAttribute(k_ans_enc_byte, cudaFuncAttributeMaxDynamicSharedMemorySize, ANS_ENC_BYTE_SMEM));
    CK(cudaFuncSetAttribute(k_dec_adaptive_seg<Leafless>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_ADAPT_SEG_SMEM));
    CK(cudaFuncSetAttribute(k_dec_adaptive_seg<LeaflessW>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_ADAPT_SEG_SMEM_WIDE));
    CK(cudaFuncSetAttribute(k_enc_adaptive2, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_AD2_SMEM));
    CK(cudaFuncSetAttribute(k_dec_static_seg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_seg_smem(true, SEG_WARPS_MAX)));
    CK(cudaFuncSetAttribute(k_dec_static_seg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_seg_smem(false, SEG_WARPS_MAX)));
    CK(cudaFuncSetAttribute(k_enc_seg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc_seg_smem(false, ENC_SEG_WARPS)));
    CK(cudaFuncSetAttribute(k_enc_seg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc_seg_smem(true, ENC_SEG_WARPS)));
    CK(cudaFuncSetAttribute(k_ans_dec_byte_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, ANS_DEC_BYTE_SEG_SMEM));
    CK(cudaFuncSetAttribute(k_enc_ranges2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_RANGES2_SMEM));
    CK(cudaFuncSetAttribute(k_enc_ranges2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_RANGES2_SMEM));
    CK(cudaFuncSetAttribute(k_blk_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, BLK_FWD_SMEM));
    CK(cudaFuncSetAttribute(k_blk_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, BLK_INV_SMEM));
    CK(cudaFuncSetAttribute(k_blk_ties, cudaFuncAttributeMaxDynamicSharedMemorySize, BLK_TIES_SMEM));
    return B2RC_OK;
}

int launch_check(b2rc_ctx* ctx, const char* what)
{
    ctx->launches += 1;
    return cuda_ok(ctx, cudaGetLastError(), what) ? B2RC_OK : B2RC_E_CUDA;
}

// Brackets one kernel launch with events when profiling is on.
struct KernelTimer {
    b2rc_ctx* ctx;
    int which;
    cudaStream_t st;
    KernelTimer(b2rc_ctx* c, int w, cudaStream_t s) : ctx(c), which(w), st(s)
    {
        if(ctx->profiling) {
            cudaEventRecord(ctx->ev[which][0], st);
        }
    }
    ~KernelTimer()
    {
        if(ctx->profiling) {
            cudaEventRecord(ctx->ev[which][1], st);
            ctx->ev_used[which] = 1;
        }
    }
};

int map_kernel_err(int bits)
{
    if(bits & ERR_CORRUPT) {
        return B2RC_E_CORRUPT;
    }
    if(bits & ERR_SLOT_OVERFLOW) {
        return B2RC_E_EXPAND;
    }
    if(bits & ERR_DST_SMALL) {
        return B2RC_E_DST_SMALL;
    }
    if(bits & ERR_INTERNAL) {
        return B2RC_E_INTERNAL;
    }
    return B2RC_OK;
}

// Pageable host memory makes every cudaMemcpyAsync a staged, synchronous copy inside the driver
// (about a quarter of the PCIe rate).  Page-locking the caller's buffer for the call (cudaHostRegister)
// was measured and is worse still (profiles/r2_ncu_notes.md).  Large pageable SOURCES are therefore
// moved chunk by chunk into pinned staging by a few threads (b2rc_host_copy) while the previous
// chunk is on the wire; B2RC_STAGE_HOST=0 switches that off.
bool is_pageable(const void* ptr)
{
    static const bool enabled = [] {
        const char* e = getenv("B2RC_STAGE_HOST");
        return !(e && atol(e) == 0);
    }();
    if(!enabled || !ptr) {
        return false;
    }
    cudaPointerAttributes a;
    if(cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(0), ok(false)
    {
        if(cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) {
            ok = true;
        }
    }
    ~DeviceGuard()
    {
        if(ok) {
            cudaSetDevice(prev);
        }
    }
};
}  // namespace

extern "C" {

const char* b2rc_build_arch(void)
{
    return "sm_100a";
}

const char* b2rc_strerror(int code)
{
    switch(code) {
    case B2RC_OK: return "ok";
    case B2RC_E_ARG: return "bad argument";
    case B2RC_E_DST_SMALL: return "destination too small";
    case B2RC_E_CORRUPT: return "corrupt container or payload";
    case B2RC_E_CUDA: return "CUDA error or no device";
    case B2RC_E_EXPAND: return "payload outgrew its staging slot";
    case B2RC_E_NOMEM: return "out of device memory";
    case B2RC_E_INTERNAL: return "internal inconsistency (a segment's bytes did not end where the range pass put them)";
    default: return "unknown";
    }
}

uint64_t b2rc_slot_bytes(uint32_t n)
{
    const u64 s = (u64)n + n / 8 + 1024;
    return (s + 127) & ~127ull;
}

uint64_t b2rc_slot_bytes_for(int mode, uint32_t n)
{
    if(is_ans(mode)) {
        // header + the flushed states + at most one 16-bit word per symbol (the word coder
        // emits one for EVERY symbol of a block that holds a single distinct byte)
        return ((u64)ANS_HDR + 32u + 2ull * n + 15u) & ~15ull;
    }
    return b2rc_slot_bytes(n);
}

uint64_t b2rc_nblocks(uint64_t n, uint32_t block_size)
{
    return block_size ? (n + block_size - 1) / block_size : 0;
}

uint64_t b2rc_bound(int mode, uint64_t n, uint32_t block_size)
{
    if(!block_ok(block_size)) {
        return 0;
    }
    const u64 nb = b2rc_nblocks(n, block_size);
    if(nb == 0) {
        return index_bytes(0);
    }
    const u64 last = n - (nb - 1) * block_size;
    // room for a restart table at the shortest segment length a context can be set to
    const u64 table = !has_restart(mode) ? 0ull
                      : 4ull + nb * 4ull * rec_words(mode, block_size) *
                                   b2rc_restart_records(block_size, mode == B2RC_MODE_ADAPTIVE ? B2RC_MIN_ADAPTIVE_RESTART_SYMS
                                                                                              : B2RC_MIN_RESTART_SYMS);
    return index_bytes(nb) + (nb - 1) * b2rc_slot_bytes_for(mode, block_size) + b2rc_slot_bytes_for(mode, (u32)last) +
           table;
}

int b2rc_ctx_create(int device, b2rc_ctx** out)
{
    if(!out) {
        return B2RC_E_ARG;
    }
    *out = nullptr;
    int count = 0;
    if(cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return B2RC_E_CUDA;  // no CPU fallback, by design
    }
    DeviceGuard g(device);
    if(!g.ok) {
        return B2RC_E_CUDA;
    }
    b2rc_ctx* ctx = new(std::nothrow) b2rc_ctx();
    if(!ctx) {
        return B2RC_E_NOMEM;
    }
    memset(ctx, 0, sizeof *ctx);
    ctx->device = device;
    // The host-pointer pipeline gives every chunk a stream of its own.  Streams beyond the number of
    // hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 unless the PROCESS exported more before its
    // first CUDA call -- a library must not edit the environment) alias, and one chunk's copy then
    // waits behind another chunk's kernel: with fewer queues, fewer chunks.
    ctx->max_chunks = B2RC_PIPE_CHUNKS;
    {
        long conn = 8;
        if(const char* e = getenv("CUDA_DEVICE_MAX_CONNECTIONS")) {
            conn = atol(e);
        }
        if(conn < B2RC_PIPE_CHUNKS + 2) {
            ctx->max_chunks = conn > 4 ? (u64)(conn - 2) : 2;
        }
    }
    ctx->min_chunk = B2RC_PIPE_MIN_CHUNK;
    if(const char* e = getenv("B2RC_PIPE_MIN_CHUNK")) {
        const long long v = atoll(e);
        if(v >= 65536) {
            ctx->min_chunk = (u64)v;
        }
    }
    ctx->ramp_chunks = 1;
    if(const char* e = getenv("B2RC_PIPE_RAMP")) {
        ctx->ramp_chunks = atol(e) ? 1u : 0u;
    }
    if(const char* e = getenv("B2RC_PIPE_CHUNKS")) {
        const long v = atol(e);
        if(v >= 1 && v <= B2RC_PIPE_CHUNKS) {
            ctx->max_chunks = (u64)v;
        }
    }
    ctx->max_phases = 4;  // measured on B200: 4 and 8 launches per chunk time the same
    ctx->seg_syms = B2RC_DEFAULT_RESTART_SYMS;
    ctx->seg_syms_static = B2RC_DEFAULT_STATIC_RESTART_SYMS;
    ctx->seg_auto = 1;
    if(const char* e = getenv("B2RC_RESTART_SYMS")) {
        const long v = atol(e);
        if(v == 0 || (v >= (long)B2RC_MIN_RESTART_SYMS && v <= (1 << 22) && v % 64 == 0)) {
            ctx->seg_syms = (u32)v;
            ctx->seg_syms_static = (u32)v;
            ctx->seg_auto = 0;  // an explicit spacing is kept for every stream
        }
    }
    ctx->enc_seg_syms = 2048;
    if(const char* e = getenv("B2RC_ENC_SEG_SYMS")) {
        const long v = atol(e);
        if(v == 0 || (v >= 64 && v <= (1 << 22) && v % 64 == 0)) {
            ctx->enc_seg_syms = (u32)v;
        }
    }
    if(const char* e = getenv("B2RC_FORCE_EXACT")) {
        ctx->force_exact = atol(e) ? 1u : 0u;
    }
    if(const char* e = getenv("B2RC_DEC_SEG_WARPS")) {
        const long v = atol(e);
        ctx->dec_seg_warps = (v >= 1 && v <= (long)SEG_WARPS_MAX) ? (u32)v : 0u;
    }
    ctx->ranges_warps = 0;  // by the size of the call (static_ranges_launch)
    if(const char* e = getenv("B2RC_RANGES_WARPS")) {
        const long v = atol(e);
        ctx->ranges_warps = (v >= 1 && v <= 3) ? (u32)v : 0u;
    }
    if(const char* e = getenv("B2RC_ADAPTIVE_TWO_WARPS")) {
        ctx->adaptive_two_warps = atol(e) ? 1u : 0u;
    }
    ctx->seg_syms_adaptive = B2RC_DEFAULT_ADAPTIVE_RESTART_SYMS;
    if(const char* e = getenv("B2RC_ADAPTIVE_RESTART_SYMS")) {
        const long v = atol(e);
        if(v == 0 || (v >= (long)B2RC_MIN_ADAPTIVE_RESTART_SYMS && v <= (1 << 22) && v % 64 == 0)) {
            ctx->seg_syms_adaptive = (u32)v;
        }
    }
    ctx->seg_syms_adaptive_wide = B2RC_DEFAULT_ADAPTIVE_RESTART_SYMS_WIDE;
    if(const char* e = getenv("B2RC_ADAPTIVE_RESTART_SYMS_WIDE")) {
        const long v = atol(e);
        if(v == 0 || (v >= (long)B2RC_MIN_ADAPTIVE_RESTART_SYMS && v <= (1 << 22) && v % 64 == 0)) {
            ctx->seg_syms_adaptive_wide = (u32)v;
        }
    }
    if(const char* e = getenv("B2RC_PHASES")) {
        const long v = atol(e);
        if(v >= 1 && v <= B2RC_PHASES) {
            ctx->max_phases = (u64)v;
        }
    }
    int rc = B2RC_OK;
    do {
        if(!cuda_ok(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate")) {
            rc = B2RC_E_CUDA;
            break;
        }
        if(!cuda_ok(ctx, cudaMalloc((void**)&ctx->d_err, 16), "cudaMalloc") ||
           !cuda_ok(ctx, cudaMalloc((void**)&ctx->d_total, 16), "cudaMalloc") ||
           !cuda_ok(ctx, cudaMallocHost((void**)&ctx->h_res, sizeof(b2rc_ctx::Result)), "cudaMallocHost")) {
            rc = B2RC_E_CUDA;
            break;
        }
        for(int k = 0; k < B2RC_PIPE_STREAMS && rc == B2RC_OK; ++k) {
            if(!cuda_ok(ctx, cudaStreamCreateWithFlags(&ctx->pipe[k], cudaStreamNonBlocking), "cudaStreamCreate")) {
                rc = B2RC_E_CUDA;
            }
        }
        for(int k = 0; k < B2RC_PIPE_CHUNKS && rc == B2RC_OK; ++k) {
            if(!cuda_ok(ctx, cudaEventCreateWithFlags(&ctx->scan_done[k], cudaEventDisableTiming), "cudaEventCreate") ||
               !cuda_ok(ctx, cudaEventCreateWithFlags(&ctx->chunk_done[k], cudaEventDisableTiming), "cudaEventCreate")) {
                rc = B2RC_E_CUDA;
            }
        }
        if(rc == B2RC_OK &&
           (!cuda_ok(ctx, cudaEventCreateWithFlags(&ctx->index_ready, cudaEventDisableTiming), "cudaEventCreate") ||
            !cuda_ok(ctx, cudaMalloc((void**)&ctx->d_ends, 8 * (B2RC_PIPE_CHUNKS + 1)), "cudaMalloc") ||
            !cuda_ok(ctx, cudaMallocHost((void**)&ctx->h_ends, 8 * (B2RC_PIPE_CHUNKS + 1)), "cudaMallocHost"))) {
            rc = B2RC_E_CUDA;
        }
        for(int k = 0; k < B2RC_PHASES && rc == B2RC_OK; ++k) {
            if(!cuda_ok(ctx, cudaStreamCreateWithFlags(&ctx->d2h[k], cudaStreamNonBlocking), "cudaStreamCreate")) {
                rc = B2RC_E_CUDA;
            }
            for(int c = 0; c < B2RC_PIPE_CHUNKS && rc == B2RC_OK; ++c) {
                if(!cuda_ok(ctx, cudaEventCreateWithFlags(&ctx->phase_done[c][k], cudaEventDisableTiming),
                            "cudaEventCreate")) {
                    rc = B2RC_E_CUDA;
                }
            }
        }
        for(int k = 0; k < B2RC_K_COUNT && rc == B2RC_OK; ++k) {
            for(int e = 0; e < 2; ++e) {
                if(!cuda_ok(ctx, cudaEventCreate(&ctx->ev[k][e]), "cudaEventCreate")) {
                    rc = B2RC_E_CUDA;
                }
            }
        }
        if(rc == B2RC_OK) {
            rc = set_smem_limits(ctx);
        }

    } while(0);
    if(rc != B2RC_OK) {
        fprintf(stderr, "b2rc_ctx_create: %s\n", ctx->last_err);
        b2rc_ctx_destroy(ctx);
        return rc;
    }
    ctx->ndev = 1;
    ctx->sub[0] = ctx;
    *out = ctx;
    return B2RC_OK;
}

int b2rc_ctx_create_multi(const int* devices, int ndev, b2rc_ctx** out)
{
    if(!out || !devices || ndev < 1 || ndev > B2RC_MAX_DEVICES) {
        return B2RC_E_ARG;
    }
    *out = nullptr;
    b2rc_ctx* head = nullptr;
    int rc = b2rc_ctx_create(devices[0], &head);
    if(rc != B2RC_OK) {
        return rc;
    }
    for(int d = 1; d < ndev; ++d) {
        b2rc_ctx* c = nullptr;
        if((rc = b2rc_ctx_create(devices[d], &c)) != B2RC_OK) {
            b2rc_ctx_destroy(head);
            return rc;
        }
        head->sub[d] = c;
        head->ndev = d + 1;
    }
    *out = head;
    return B2RC_OK;
}

uint32_t b2rc_restart_for(const b2rc_ctx* ctx, int mode, uint32_t block_size, uint64_t nblocks)
{
    return (ctx && mode_ok(mode) && block_ok(block_size)) ? seg_for(ctx, mode, block_size, nblocks) : 0u;
}

int b2rc_ctx_force_restart(b2rc_ctx* ctx, uint32_t seg_syms)
{
    if(!ctx || (seg_syms && (seg_syms < B2RC_MIN_RESTART_SYMS || seg_syms > (1u << 22) || seg_syms % 64u))) {
        return B2RC_E_ARG;
    }
    ctx->seg_force = seg_syms;
    for(int d = 0; d < ctx->ndev && ctx->ndev > 1; ++d) {
        if(ctx->sub[d] && ctx->sub[d] != ctx) {
            ctx->sub[d]->seg_force = seg_syms;
        }
    }
    return B2RC_OK;
}

int b2rc_ctx_devices(const b2rc_ctx* ctx)
{
    return ctx ? ctx->ndev : 0;
}

void b2rc_ctx_destroy(b2rc_ctx* ctx)
{
    if(!ctx) {
        return;
    }
    for(int d = 1; d < ctx->ndev; ++d) {
        b2rc_ctx_destroy(ctx->sub[d]);
        ctx->sub[d] = nullptr;
    }
    ctx->ndev = 0;
    DeviceGuard g(ctx->device);
    if(ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(ctx->slots);
    cudaFree(ctx->sizes);
    cudaFree(ctx->freq16);
    cudaFree(ctx->stage_in);
    cudaFree(ctx->stage_out);
    cudaFree(ctx->d_err);
    cudaFree(ctx->blk_rounds);
    cudaFree(ctx->blk_tie_list);
    cudaFree(ctx->blk_tmp);
    cudaFree(ctx->blk_ties);
    cudaFree(ctx->d_total);
    if(ctx->h_res) {
        cudaFreeHost(ctx->h_res);
    }
    if(ctx->h_stage) {
        cudaFreeHost(ctx->h_stage);
    }
    if(ctx->h_in) {
        cudaFreeHost(ctx->h_in);
    }
    for(int k = 0; k < B2RC_PIPE_STREAMS; ++k) {
        if(ctx->pipe[k]) {
            cudaStreamSynchronize(ctx->pipe[k]);
            cudaStreamDestroy(ctx->pipe[k]);
        }
    }
    for(int k = 0; k < B2RC_PIPE_CHUNKS; ++k) {
        if(ctx->scan_done[k]) {
            cudaEventDestroy(ctx->scan_done[k]);
        }
        if(ctx->chunk_done[k]) {
            cudaEventDestroy(ctx->chunk_done[k]);
        }
    }
    if(ctx->index_ready) {
        cudaEventDestroy(ctx->index_ready);
    }
    for(int k = 0; k < B2RC_PHASES; ++k) {
        if(ctx->d2h[k]) {
            cudaStreamSynchronize(ctx->d2h[k]);
            cudaStreamDestroy(ctx->d2h[k]);
        }
        for(int c = 0; c < B2RC_PIPE_CHUNKS; ++c) {
            if(ctx->phase_done[c][k]) {
                cudaEventDestroy(ctx->phase_done[c][k]);
            }
        }
    }
    cudaFree(ctx->dec_state);
    cudaFree(ctx->restart);
    cudaFree(ctx->seg_recs);
    cudaFree(ctx->seg_lows);
    cudaFree(ctx->dec_model);
    cudaFree(ctx->d_ends);
    if(ctx->h_ends) {
        cudaFreeHost(ctx->h_ends);
    }
    for(int k = 0; k < B2RC_K_COUNT; ++k) {
        for(int e = 0; e < 2; ++e) {
            if(ctx->ev[k][e]) {
                cudaEventDestroy(ctx->ev[k][e]);
            }
        }
    }
    if(ctx->stream) {
        cudaStreamDestroy(ctx->stream);
    }
    cudaGetLastError();
    delete ctx;
}

const char* b2rc_last_cuda_error(const b2rc_ctx* ctx)
{
    return ctx ? ctx->last_err : "";
}

uint64_t b2rc_launch_count(const b2rc_ctx* ctx)
{
    return ctx ? ctx->launches : 0;
}

int b2rc_profile(b2rc_ctx* ctx, int enable)
{
    if(!ctx) {
        return B2RC_E_ARG;
    }
    ctx->profiling = enable ? 1 : 0;
    memset(ctx->ev_used, 0, sizeof ctx->ev_used);
    return B2RC_OK;
}

int b2rc_kernel_ms(b2rc_ctx* ctx, int which, float* ms)
{
    if(!ctx || !ms || which < 0 || which >= B2RC_K_COUNT) {
        return B2RC_E_ARG;
    }
    *ms = -1.0f;
    if(!ctx->ev_used[which]) {
        return B2RC_OK;
    }
    DeviceGuard g(ctx->device);
    CK(cudaEventSynchronize(ctx->ev[which][1]));
    CK(cudaEventElapsedTime(ms, ctx->ev[which][0], ctx->ev[which][1]));
    return B2RC_OK;
}

int b2rc_peek(const uint8_t* src, uint64_t n, int* mode, uint32_t* block_size, uint64_t* total, uint64_t* nblocks)
{
    if(!src) {
        return B2RC_E_ARG;
    }
    if(n < B2RC_HEADER_BYTES + 8) {
        return B2RC_E_CORRUPT;
    }
    u32 h[8];
    memcpy(h, src, sizeof h);
    const u32 version = h[1] & 0xFFFFu, md = h[1] >> 16;
    const u64 tot = (u64)h[4] | ((u64)h[5] << 32), nb = (u64)h[6] | ((u64)h[7] << 32);
    if(h[0] != 0x43523242u || version != 1u || !mode_ok((int)md) || !block_ok(h[2])) {
        return B2RC_E_CORRUPT;
    }
    if(h[3] != 0u) {  // a restart table: static range coder or byte rANS, a legal segment length, nothing else set
        const u32 seg = (h[3] >> 8) * 64u;
        if((h[3] & 0xFFu) != 1u || !has_restart((int)md) || !seg_ok(h[2], seg)) {
            return B2RC_E_CORRUPT;
        }
    }
    if(nb != b2rc_nblocks(tot, h[2]) || nb > (n - B2RC_HEADER_BYTES) / 8 - 1) {
        return B2RC_E_CORRUPT;
    }
    if(mode) {
        *mode = (int)md;
    }
    if(block_size) {
        *block_size = h[2];
    }
    if(total) {
        *total = tot;
    }
    if(nblocks) {
        *nblocks = nb;
    }
    return B2RC_OK;
}

// ------------------------------------------------------------ rANS launches --
static int ans_encode_blocks(b2rc_ctx* ctx, int mode, u32 block, const u8* d_src, u64 n, u8* d_slots, u64 stride,
                             u32* d_sizes, u32* d_restart, u32 seg_syms, int* d_err, cudaStream_t st)
{
    const u64 nb = b2rc_nblocks(n, block);
    if(nb == 0) {
        return B2RC_OK;
    }
    DeviceGuard g(ctx->device);
    u64 grid = (nb + HIST_WARPS - 1) / HIST_WARPS;
    if(grid > 148ull * 16) {
        grid = 148ull * 16;
    }
    {
        KernelTimer kt(ctx, B2RC_K_HISTOGRAM, st);
        if(mode == B2RC_MODE_RANS_WORD) {
            k_ans_model<(int)ANS_WORD_BITS><<<(unsigned)grid, HIST_WARPS * 32, 0, st>>>(d_src, n, block, nb, d_slots,
                                                                                      stride);
        } else {
            k_ans_model<(int)ANS_BYTE_BITS><<<(unsigned)grid, HIST_WARPS * 32, 0, st>>>(d_src, n, block, nb, d_slots,
                                                                                      stride);
        }
        const int rc = launch_check(ctx, "k_ans_model");
        if(rc != B2RC_OK) {
            return rc;
        }
    }
    EncArgs a;
    a.src = d_src;
    a.n = n;
    a.block = block;
    a.nblocks = nb;
    a.freq16 = nullptr;
    a.slots = d_slots;
    a.slot_stride = stride;
    a.sizes = d_sizes;
    a.err = d_err;
    a.restart = mode == B2RC_MODE_RANS_BYTE ? d_restart : nullptr;
    a.seg_syms = a.restart ? seg_syms : 0u;
    if(a.restart) {
        CK(cudaMemsetAsync(d_restart, 0xFF, (size_t)(nb * b2rc_restart_records(block, seg_syms) * 12u), st));
    }
    KernelTimer kt(ctx, B2RC_K_ENCODE, st);
    if(mode == B2RC_MODE_RANS_WORD) {
        k_ans_enc_word<<<(unsigned)((nb + 3) / 4), 32, 0, st>>>(a);
    } else {
        k_ans_enc_byte<<<(unsigned)((nb + 31) / 32), 32, ANS_ENC_BYTE_SMEM, st>>>(a);
    }
    return launch_check(ctx, "k_ans_enc");
}

static int ans_decode_blocks(b2rc_ctx* ctx, int mode, const DecArgs& a, cudaStream_t st)
{
    KernelTimer kt(ctx, B2RC_K_DECODE, st);
    if(mode == B2RC_MODE_RANS_BYTE) {
        k_ans_dec_byte<<<(unsigned)((a.nblocks + 31) / 32), 32, ANS_DEC_BYTE_SMEM, st>>>(a);
        return launch_check(ctx, "k_ans_dec_byte");
    }
    k_ans_dec_word<<<(unsigned)((a.nblocks + 4 * ANS_DEC_WARPS - 1) / (4 * ANS_DEC_WARPS)), 32 * ANS_DEC_WARPS,
                     ANS_DEC_WORD_SMEM, st>>>(a);
    return launch_check(ctx, "k_ans_dec_word");
}

// ------------------------------------------------------------ kernel doors --
int b2rc_k_histogram(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint32_t block_size, uint16_t* d_freq16,
                     void* cuda_stream)
{
    if(!ctx || (n && (!d_src || !d_freq16)) || !block_ok(block_size) || !aligned16(d_src) || !aligned16(d_freq16)) {
        return B2RC_E_ARG;  // an empty range (a rank without blocks) may come with null pointers
    }
    const u64 nb = b2rc_nblocks(n, block_size);
    if(nb == 0) {
        return B2RC_OK;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    u64 grid = (nb + HIST_WARPS - 1) / HIST_WARPS;
    if(grid > 148ull * 16) {
        grid = 148ull * 16;
    }
    KernelTimer kt(ctx, B2RC_K_HISTOGRAM, st);
    if(block_size <= 65536u) {
        k_hist<<<(unsigned)grid, HIST_WARPS * 32, 0, st>>>(d_src, n, block_size, nb, d_freq16);
    } else {
        k_hist_wide<<<(unsigned)grid, HIST_WARPS * 32, 0, st>>>(d_src, n, block_size, nb, d_freq16);
    }
    return launch_check(ctx, "k_hist");
}


uint32_t b2rc_restart_records(uint32_t block_size, uint32_t seg_syms)
{
    return seg_ok(block_size, seg_syms) ? (block_size + seg_syms - 1u) / seg_syms - 1u : 0u;
}

int b2rc_k_encode_blocks(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                         const uint16_t* d_freq16, uint8_t* d_slots, uint64_t slot_stride, uint32_t* d_sizes,
                         int* d_err, void* cuda_stream)
{
    return b2rc_k_encode_blocks_r(ctx, mode, block_size, d_src, n, d_freq16, d_slots, slot_stride, d_sizes, nullptr, 0u,
                                  d_err, cuda_stream);
}

int b2rc_k_encode_blocks_r(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                           const uint16_t* d_freq16, uint8_t* d_slots, uint64_t slot_stride, uint32_t* d_sizes,
                           uint32_t* d_restart, uint32_t seg_syms, int* d_err, void* cuda_stream)
{
    if(d_restart && (!has_restart(mode) || !seg_ok(block_size, seg_syms) || ((uintptr_t)d_restart & 3u))) {
        return B2RC_E_ARG;
    }
    if(ctx && n == 0 && mode_ok(mode) && block_ok(block_size)) {
        return B2RC_OK;  // nothing to code: a rank without blocks (world > nblocks) passes null pointers
    }
    if(!ctx || !d_src || !d_slots || !d_sizes || !d_err || !mode_ok(mode) || !block_ok(block_size) ||
       !aligned16(d_src) || !aligned16(d_slots) || (slot_stride & 15u) ||
       slot_stride < b2rc_slot_bytes_for(mode, block_size) || slot_stride > 0xFFFFFFF0ull) {
        return B2RC_E_ARG;
    }
    if(is_ans(mode)) {
        return ans_encode_blocks(ctx, mode, block_size, d_src, n, d_slots, slot_stride, d_sizes, d_restart, seg_syms, d_err,
                                 (cudaStream_t)cuda_stream);
    }
    const bool wide = block_size > 65536u;
    if(mode == B2RC_MODE_STATIC && ((!wide && !d_freq16) || (d_freq16 && !aligned16(d_freq16)))) {
        return B2RC_E_ARG;
    }
    const u64 nb = b2rc_nblocks(n, block_size);
    if(nb == 0) {
        return B2RC_OK;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    EncArgs a;
    a.src = d_src;
    a.n = n;
    a.block = block_size;
    a.nblocks = nb;
    a.freq16 = d_freq16;
    a.slots = d_slots;
    a.slot_stride = slot_stride;
    a.sizes = d_sizes;
    a.err = d_err;
    a.restart = d_restart;
    a.seg_syms = d_restart ? seg_syms : 0u;
    if(d_restart) {
        // 0xFF: "no such point" for blocks shorter than the segment start
        CK(cudaMemsetAsync(d_restart, 0xFF, (size_t)(nb * b2rc_restart_records(block_size, seg_syms) * 4u * rec_words(mode, block_size)), st));
    }
    const unsigned grid = (unsigned)((nb + 31) / 32);
    KernelTimer kt(ctx, B2RC_K_ENCODE, st);
    if(mode == B2RC_MODE_STATIC) {
        if(wide) {
            k_enc_static<true><<<grid, 32, smem_enc_static(true), st>>>(a);
        } else {
            k_enc_static<false><<<grid, 32, smem_enc_static(false), st>>>(a);
        }
    } else {
        if(wide) {
            k_enc_adaptive<u32><<<grid, 32, smem_enc_adaptive(true), st>>>(a);
        } else if(ctx->adaptive_two_warps) {
            k_enc_adaptive2<<<grid, 64, ENC_AD2_SMEM, st>>>(a);  // model and coder on a warp each: measured slower, kept for the record
        } else {
            k_enc_adaptive<u16><<<grid, 32, smem_enc_adaptive(false), st>>>(a);
        }
    }
    return launch_check(ctx, "k_enc");
}

static int scan_launch(b2rc_ctx* ctx, const u32* d_sizes, u64 nb, u64* d_offsets, u64* d_total, u8* d_header, u32 mode,
                       u32 block, u64 n, cudaStream_t st, const u64* d_base = nullptr, u32 flags = 0)
{
    KernelTimer kt(ctx, B2RC_K_SCAN, st);
    k_scan<<<1, SCAN_THREADS, 0, st>>>(d_sizes, nb, d_offsets, d_total, d_header, mode, block, n, d_base, flags);
    return launch_check(ctx, "k_scan");
}

int b2rc_k_scan(b2rc_ctx* ctx, const uint32_t* d_sizes, uint64_t nblocks, uint64_t* d_offsets, void* cuda_stream)
{
    if(!ctx || !d_offsets || (nblocks && !d_sizes) || ((uintptr_t)d_offsets & 7u)) {
        return B2RC_E_ARG;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    return scan_launch(ctx, d_sizes, nblocks, d_offsets, nullptr, nullptr, 0, 0, 0, st);
}

int b2rc_k_compact(b2rc_ctx* ctx, const uint8_t* d_slots, uint64_t slot_stride, const uint32_t* d_sizes,
                   const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_payload, uint64_t payload_cap, int* d_err,
                   void* cuda_stream)
{
    return b2rc_k_compact_for(ctx, B2RC_MODE_STATIC, d_slots, slot_stride, d_sizes, d_offsets, nblocks, d_payload,
                              payload_cap, d_err, cuda_stream);
}

int b2rc_k_compact_for(b2rc_ctx* ctx, int mode, const uint8_t* d_slots, uint64_t slot_stride, const uint32_t* d_sizes,
                       const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_payload, uint64_t payload_cap,
                       int* d_err, void* cuda_stream)
{
    if(!ctx || !mode_ok(mode) || !d_slots || !d_sizes || !d_offsets || !d_payload || !d_err || !aligned16(d_slots) ||
       (slot_stride & 15u)) {
        return B2RC_E_ARG;
    }
    if(nblocks == 0) {
        return B2RC_OK;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    u64 grid = nblocks;
    if(grid > 148ull * 8) {
        grid = 148ull * 8;
    }
    KernelTimer kt(ctx, B2RC_K_COMPACT, st);
    if(is_ans(mode)) {
        k_compact_split<<<(unsigned)grid, COMPACT_THREADS, 0, st>>>(d_slots, slot_stride, d_sizes, d_offsets, nblocks,
                                                                    d_payload, payload_cap, d_err, ANS_HDR);
    } else {
        k_compact<<<(unsigned)grid, COMPACT_THREADS, 0, st>>>(d_slots, slot_stride, d_sizes, d_offsets, nblocks,
                                                              d_payload, payload_cap, d_err);
    }
    return launch_check(ctx, "k_compact");
}

static int decode_launch(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_payload, uint64_t payload_len,
                         const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_dst, uint64_t n, int* d_err,
                         void* cuda_stream, u32 sym0, u32 nsym, u32* d_state, u8* d_model);

int b2rc_k_decode_blocks(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_payload,
                         uint64_t payload_len, const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_dst, uint64_t n,
                         int* d_err, void* cuda_stream)
{
    return decode_launch(ctx, mode, block_size, d_payload, payload_len, d_offsets, nblocks, d_dst, n, d_err, cuda_stream,
                         0u, 0u, nullptr, nullptr);
}

int b2rc_k_decode_blocks_r(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_payload,
                           uint64_t payload_len, const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_dst, uint64_t n,
                           const uint32_t* d_restart, uint32_t seg_syms, int* d_err, void* cuda_stream)
{
    if(!d_restart) {
        return b2rc_k_decode_blocks(ctx, mode, block_size, d_payload, payload_len, d_offsets, nblocks, d_dst, n, d_err,
                                    cuda_stream);
    }
    if(!ctx || !d_payload || !d_offsets || !d_dst || !d_err || !has_restart(mode) || !block_ok(block_size) ||
       !seg_ok(block_size, seg_syms) || !aligned16(d_dst) || ((uintptr_t)d_offsets & 7u) || ((uintptr_t)d_restart & 3u) ||
       nblocks != b2rc_nblocks(n, block_size)) {
        return B2RC_E_ARG;
    }
    if(nblocks == 0) {
        return B2RC_OK;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DecArgs a;
    a.payload = d_payload;
    a.payload_len = payload_len;
    a.offsets = d_offsets;
    a.nblocks = nblocks;
    a.block = block_size;
    a.dst = d_dst;
    a.n = n;
    a.err = d_err;
    a.sym0 = 0;
    a.nsym = 0;
    a.state = nullptr;
    a.model = nullptr;
    a.restart = d_restart;
    a.seg_syms = seg_syms;
    const u32 nseg = b2rc_restart_records(block_size, seg_syms) + 1u;
    KernelTimer kt(ctx, B2RC_K_DECODE, st);
    if(mode == B2RC_MODE_ADAPTIVE) {
        const dim3 agrid((unsigned)((nblocks + 31) / 32), nseg);
        if(block_size <= 65536u) {
            k_dec_adaptive_seg<Leafless><<<agrid, 32, DEC_ADAPT_SEG_SMEM, st>>>(a);
        } else {
            k_dec_adaptive_seg<LeaflessW><<<agrid, 32, DEC_ADAPT_SEG_SMEM_WIDE, st>>>(a);
        }
        return launch_check(ctx, "k_dec_adaptive_seg");
    }
    if(mode == B2RC_MODE_RANS_BYTE) {
        const dim3 agrid((unsigned)((nblocks + 31) / 32), (nseg + ANS_SEG_WARPS - 1u) / ANS_SEG_WARPS);
        k_ans_dec_byte_seg<<<agrid, 32 * ANS_SEG_WARPS, ANS_DEC_BYTE_SEG_SMEM, st>>>(a);
        return launch_check(ctx, "k_ans_dec_byte_seg");
    }
    // sixteen or more segments per block: CTAs of eight warps (one table per eight chains, 32 warps per SM instead
    // of 28); the eight segments of the older spacing stay two CTAs of four (eight in one leave some SMs with 24
    // warps and others with 32: 4.40 against 4.04 ms)
    const u32 warps = ctx->dec_seg_warps ? ctx->dec_seg_warps : (nseg >= 16u ? 8u : (nseg < SEG_WARPS ? nseg : SEG_WARPS));
    const dim3 grid((unsigned)((nblocks + 31) / 32), (nseg + warps - 1u) / warps);
    if(block_size <= 65536u) {
        k_dec_static_seg<true><<<grid, 32 * warps, dec_seg_smem(true, warps), st>>>(a);
    } else {
        k_dec_static_seg<false><<<grid, 32 * warps, dec_seg_smem(false, warps), st>>>(a);
    }
    return launch_check(ctx, "k_dec_static_seg");
}

// sym0 / nsym / d_state / d_model: phased decode of the range coders (DecArgs); 0, 0, null, null =
// whole blocks
static int decode_launch(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_payload, uint64_t payload_len,
                         const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_dst, uint64_t n, int* d_err,
                         void* cuda_stream, u32 sym0, u32 nsym, u32* d_state, u8* d_model)
{
    if(!ctx || !d_payload || !d_offsets || !d_dst || !d_err || !mode_ok(mode) || !block_ok(block_size) ||
       !aligned16(d_dst) || ((uintptr_t)d_offsets & 7u) || nblocks != b2rc_nblocks(n, block_size) ||
       (mode == B2RC_MODE_RANS_WORD && ((uintptr_t)d_payload & 1u))) {
        return B2RC_E_ARG;
    }
    if(nblocks == 0) {
        return B2RC_OK;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DecArgs a;
    a.payload = d_payload;
    a.payload_len = payload_len;
    a.offsets = d_offsets;
    a.nblocks = nblocks;
    a.block = block_size;
    a.dst = d_dst;
    a.n = n;
    a.err = d_err;
    a.sym0 = sym0;
    a.nsym = nsym;
    a.state = d_state;
    a.model = d_model;
    a.restart = nullptr;
    a.seg_syms = 0;
    if(nsym && (is_ans(mode) || !d_state || (mode == B2RC_MODE_ADAPTIVE && !d_model) || (sym0 % TILE) || (nsym % TILE))) {
        return B2RC_E_ARG;
    }
    if(is_ans(mode)) {
        return ans_decode_blocks(ctx, mode, a, st);
    }
    const bool wide = block_size > 65536u;
    const unsigned grid = (unsigned)((nblocks + 31) / 32);
    KernelTimer kt(ctx, B2RC_K_DECODE, st);
    if(mode == B2RC_MODE_STATIC) {
        if(nsym) {
            k_dec_static<true><<<grid, 32, smem_dec_static(), st>>>(a);
        } else {
            k_dec_static<false><<<grid, 32, smem_dec_static(), st>>>(a);
        }
    } else {
        if(wide) {
            if(nsym) {
                k_dec_adaptive<u32, true><<<grid, 32, smem_dec_adaptive(true), st>>>(a);
            } else {
                k_dec_adaptive<u32, false><<<grid, 32, smem_dec_adaptive(true), st>>>(a);
            }
        } else {
            if(nsym) {
                k_dec_adaptive<u16, true><<<grid, 32, smem_dec_adaptive(false), st>>>(a);
            } else {
                k_dec_adaptive<u16, false><<<grid, 32, smem_dec_adaptive(false), st>>>(a);
            }
        }
    }
    return launch_check(ctx, "k_dec");
}


// ------------------------------------------- segmented static encode (b2rc_encseg.cuh) --
// Symbols per segment for this block size; it must divide the restart spacing so that every
// restart point is a segment start.  0: this context codes one chain per block (k_enc_static).
static u32 enc_seg_plan(const b2rc_ctx* ctx, u32 block, u32 restart_seg)
{
    u32 P = ctx->enc_seg_syms;
    if(P == 0) {
        return 0;
    }
    if(restart_seg) {
        while(restart_seg % P) {
            P -= 64;
        }
    }
    return P < block ? P : block;
}

static int seg_scratch(b2rc_ctx* ctx, u64 nb, u32 block, u32 P)
{
    const u64 nseg = (block + P - 1) / P;
    int rc;
    if((rc = grow(ctx, ctx->seg_recs, ctx->seg_recs_cap, (size_t)(nb * (nseg + 1) * 8 + 16))) != B2RC_OK ||
       (rc = grow(ctx, ctx->seg_lows, ctx->seg_lows_cap, (size_t)(nb * nseg * 4 + 16))) != B2RC_OK) {
        return rc;
    }
    return B2RC_OK;
}

// `first` = index of the range's first block in the context's scratch (chunked host pipeline)
static SegArgs seg_args(b2rc_ctx* ctx, u32 block, u32 P, const u8* d_src, u64 n, const u16* freq, u64 first, u32* d_sizes,
                        const u64* d_offsets, u8* d_payload, u64 payload_cap, u32* d_restart, u32 seg)
{
    SegArgs a;
    a.src = d_src;
    a.n = n;
    a.block = block;
    a.nblocks = b2rc_nblocks(n, block);
    a.freq16 = freq;
    a.P = P;
    a.nseg = (block + P - 1) / P;
    a.recs = ctx->seg_recs + first * (u64)(a.nseg + 1) * 2;
    a.lows = ctx->seg_lows + first * (u64)a.nseg;
    a.sizes = d_sizes;
    a.offsets = d_offsets;
    a.payload = d_payload;
    a.payload_cap = payload_cap;
    a.restart = d_restart;
    a.seg_syms = d_restart ? seg : 0u;
    a.err = ctx->d_err;
    a.force_exact = ctx->force_exact;
    return a;
}

// K2r: sizes of all payloads (and where every segment starts), before a byte is coded
// `call_blocks`: blocks of the whole call (the chunks of the host pipeline run side by side).  A stream that leaves
// most of the GPU idle -- at most one CTA per SM -- takes the three-warp kernel for 64 KiB blocks: 0.75 instead of
// 1.06 ms; with more CTAs than SMs the extra warps get in each other's way (1 GiB: 2.1 against 1.08 ms).
static int static_ranges_launch(b2rc_ctx* ctx, const SegArgs& a, u64 call_blocks, cudaStream_t st)
{
    const unsigned grid = (unsigned)((a.nblocks + 31) / 32);
    const u32 warps = ctx->ranges_warps ? ctx->ranges_warps : ((call_blocks + 31) / 32 <= 148u ? 3u : 1u);
    KernelTimer kt(ctx, B2RC_K_RANGES, st);
    if(a.block > 65536u) {
        k_enc_ranges<true><<<grid, 32, ENC_RANGES_SMEM, st>>>(a);
    } else if(a.block == 65536u && warps == 3) {  // total 65536 happens for these only (b2rc_encseg.cuh)
        k_enc_ranges2<3><<<grid, 96, ENC_RANGES2_SMEM, st>>>(a);
    } else if(a.block == 65536u && warps == 2) {
        k_enc_ranges2<2><<<grid, 64, ENC_RANGES2_SMEM, st>>>(a);
    } else {
        k_enc_ranges<false><<<grid, 32, ENC_RANGES_SMEM, st>>>(a);
    }
    return launch_check(ctx, "k_enc_ranges");
}

// K2s + K2m: the segments into their final place, then the seams
static int static_segments_launch(b2rc_ctx* ctx, const SegArgs& a, cudaStream_t st)
{
    const u32 warps = a.nseg < ENC_SEG_WARPS ? a.nseg : ENC_SEG_WARPS;
    const dim3 grid((unsigned)((a.nblocks + 31) / 32), (a.nseg + warps - 1) / warps);
    const bool wide = a.block > 65536u;
    {
        KernelTimer kt(ctx, B2RC_K_ENCODE, st);
        if(wide) {
            k_enc_seg<true><<<grid, 32 * warps, enc_seg_smem(true, warps), st>>>(a);
        } else {
            k_enc_seg<false><<<grid, 32 * warps, enc_seg_smem(false, warps), st>>>(a);
        }
        const int rc = launch_check(ctx, "k_enc_seg");
        if(rc != B2RC_OK) {
            return rc;
        }
    }
    KernelTimer kt(ctx, B2RC_K_SEAMS, st);
    k_enc_seams<<<(unsigned)((a.nblocks + SEAM_THREADS - 1) / SEAM_THREADS), SEAM_THREADS, 0, st>>>(a);
    return launch_check(ctx, "k_enc_seams");
}

// ------------------------------------------------------ container, device --
int b2rc_encode_device(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                       uint8_t* d_dst, uint64_t dst_cap, uint64_t* out_n, void* cuda_stream)
{
    if(!ctx || !d_dst || (n && !d_src) || !mode_ok(mode) || !block_ok(block_size) || !aligned16(d_src) ||
       !aligned16(d_dst)) {
        return B2RC_E_ARG;
    }
    const u64 nb = b2rc_nblocks(n, block_size);
    const u64 idx = index_bytes(nb);
    if(out_n) {
        *out_n = b2rc_bound(mode, n, block_size);
    }
    if(dst_cap < idx) {
        return B2RC_E_DST_SMALL;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const u64 stride = b2rc_slot_bytes_for(mode, block_size);
    const u32 seg = nb ? seg_for(ctx, mode, block_size, nb) : 0u;
    // static coder: many chains per block, coded straight into the container (b2rc_encseg.cuh)
    const u32 P = (mode == B2RC_MODE_STATIC && nb) ? enc_seg_plan(ctx, block_size, seg) : 0u;
    int rc;
    if((!P && (rc = grow(ctx, ctx->slots, ctx->slots_cap, (size_t)(nb * stride))) != B2RC_OK) ||
       (P && (rc = seg_scratch(ctx, nb, block_size, P)) != B2RC_OK) ||
       (rc = grow(ctx, ctx->sizes, ctx->sizes_cap, (size_t)(nb * 4 + 16))) != B2RC_OK) {
        return rc;
    }
    const bool need_hist = mode == B2RC_MODE_STATIC;
    if(need_hist && (rc = grow(ctx, ctx->freq16, ctx->freq_cap, (size_t)(nb * 512 + 16))) != B2RC_OK) {
        return rc;
    }
    const u64 table_words = seg ? nb * (u64)rec_words(mode, block_size) * b2rc_restart_records(block_size, seg) : 0ull;
    if(seg && (rc = grow(ctx, ctx->restart, ctx->restart_cap, (size_t)(table_words * 4 + 16))) != B2RC_OK) {
        return rc;
    }
    CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), st));
    u64* d_offsets = reinterpret_cast<u64*>(d_dst + B2RC_HEADER_BYTES);
    if(P) {
        const SegArgs a = seg_args(ctx, block_size, P, d_src, n, ctx->freq16, 0, ctx->sizes, d_offsets, d_dst + idx,
                                   dst_cap - idx, seg ? ctx->restart : nullptr, seg);
        if((rc = b2rc_k_histogram(ctx, d_src, n, block_size, ctx->freq16, st)) != B2RC_OK ||
           (rc = static_ranges_launch(ctx, a, nb, st)) != B2RC_OK ||
           (rc = scan_launch(ctx, ctx->sizes, nb, d_offsets, ctx->d_total, d_dst, (u32)mode, block_size, n, st, nullptr,
                             flags_of(seg))) != B2RC_OK ||
           (rc = static_segments_launch(ctx, a, st)) != B2RC_OK) {
            return rc;
        }
    } else {
        if(nb) {
            if(need_hist && (rc = b2rc_k_histogram(ctx, d_src, n, block_size, ctx->freq16, st)) != B2RC_OK) {
                return rc;
            }
            if((rc = b2rc_k_encode_blocks_r(ctx, mode, block_size, d_src, n, need_hist ? ctx->freq16 : nullptr, ctx->slots,
                                            stride, ctx->sizes, seg ? ctx->restart : nullptr, seg, ctx->d_err, st)) !=
               B2RC_OK) {
                return rc;
            }
        }
        if((rc = scan_launch(ctx, ctx->sizes, nb, d_offsets, ctx->d_total, d_dst, (u32)mode, block_size, n, st, nullptr,
                             flags_of(seg))) != B2RC_OK) {
            return rc;
        }
        if(nb && (rc = b2rc_k_compact_for(ctx, mode, ctx->slots, stride, ctx->sizes, d_offsets, nb, d_dst + idx,
                                          dst_cap - idx, ctx->d_err, st)) != B2RC_OK) {
            return rc;
        }
    }
    if(seg) {
        k_put_table<<<148, 256, 0, st>>>(ctx->restart, table_words, d_dst + idx, ctx->d_total, dst_cap - idx, ctx->d_err);
        if((rc = launch_check(ctx, "k_put_table")) != B2RC_OK) {
            return rc;
        }
    }
    CK(cudaMemcpyAsync(&ctx->h_res->err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&ctx->h_res->total, ctx->d_total, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if(out_n) {
        *out_n = idx + (seg ? align4(ctx->h_res->total) + 4ull * table_words : ctx->h_res->total);
    }
    return map_kernel_err(ctx->h_res->err);
}

int b2rc_decode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                       uint64_t* out_n, void* cuda_stream)
{
    if(!ctx || !d_src || !aligned16(d_src) || !aligned16(d_dst)) {
        return B2RC_E_ARG;
    }
    if(n < B2RC_HEADER_BYTES + 8) {
        return B2RC_E_CORRUPT;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CK(cudaMemcpyAsync(ctx->h_res->header, d_src, B2RC_HEADER_BYTES, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    u32 flags;
    memcpy(&flags, ctx->h_res->header + 12, 4);
    int mode;
    u32 block;
    u64 total, nb;
    // the offsets live on the device; peek only needs the header and the container length
    u8 probe[B2RC_HEADER_BYTES + 8];
    memcpy(probe, ctx->h_res->header, B2RC_HEADER_BYTES);
    memset(probe + B2RC_HEADER_BYTES, 0, 8);
    int rc = b2rc_peek(probe, n, &mode, &block, &total, &nb);
    if(rc != B2RC_OK) {
        return rc;
    }
    if(out_n) {
        *out_n = total;
    }
    if(dst_cap < total || (total && !d_dst)) {
        return B2RC_E_DST_SMALL;
    }
    if(nb == 0) {
        return B2RC_OK;
    }
    const u64 idx = index_bytes(nb);
    const u32* d_restart = nullptr;
    u32 seg = 0;
    if(flags) {
        // the table sits behind the payloads: one more small read to learn where they end
        seg = (flags >> 8) * 64u;
        CK(cudaMemcpyAsync(&ctx->h_res->total, d_src + B2RC_HEADER_BYTES + 8 * nb, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const u64 pay = ctx->h_res->total;
        const u64 bytes = nb * 4ull * rec_words(mode, block) * b2rc_restart_records(block, seg);
        if(pay > n - idx || align4(pay) + bytes > n - idx) {
            return B2RC_E_CORRUPT;
        }
        d_restart = reinterpret_cast<const u32*>(d_src + idx + align4(pay));
    }
    CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), st));
    // payload_len bounds every offset the kernel reads (dec_setup)
    if((rc = b2rc_k_decode_blocks_r(ctx, mode, block, d_src + idx, n - idx,
                                    reinterpret_cast<const u64*>(d_src + B2RC_HEADER_BYTES), nb, d_dst, total, d_restart,
                                    seg, ctx->d_err, st)) != B2RC_OK) {
        return rc;
    }
    CK(cudaMemcpyAsync(&ctx->h_res->err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return map_kernel_err(ctx->h_res->err);
}

// -------------------------------------------------------- container, host --
// The host-pointer calls pipeline: the input is cut into chunks of whole blocks; each
// chunk's H2D copy, kernels and D2H copy go to a stream of its own, so chunk c+1 is on
// the wire while chunk c is being coded and chunk c-1 is on its way back.  The coder
// kernels are latency bound (a launch takes as long for 64 MiB as for 1 GiB) and leave
// most of the GPU idle, so the kernels of all chunks overlap; what must not happen is two
// chunks queueing behind each other on one stream.  With pageable host memory the copies serialise inside the
// driver; give pinned buffers for full overlap (bench.py does).
namespace
{
struct Chunks {
    u64 nb, count;
    u64 bnd[B2RC_PIPE_CHUNKS + 1];  // chunk c = blocks [bnd[c], bnd[c + 1])
    u64 lo(u64 c) const { return bnd[c]; }
    u64 hi(u64 c) const { return bnd[c + 1]; }
};
// The pipeline's time is the longer direction's copies plus what cannot overlap them: the first chunk's way in and
// its kernels (decode: the copies home are the longer direction) and the last chunk's way home (encode: the
// copies in are).  So the chunks at both ends are shorter, growing by about 1.4 a step -- the next chunk in has to
// be there when the previous one has gone home, or the wire idles.
// (Cutting b2rc_encode's input into 32 chunks instead of 16 was tried -- copies alone in the pipeline's shape get
// faster with more chunks one way (1 GiB in, 0.79 GiB home: 23.5 / 22.5 / 21.8 ms for 16 / 32 / 64) and slower the
// other (23.8 / 24.1 / 24.8), tools/e2e_phases.py -- and changed nothing once the kernels sit between the copies.)
// `coarse`: a pageable source is moved into pinned staging chunk by chunk by the host's threads (b2rc_host_copy), and
// that copy is the call's bottleneck: six long chunks keep every thread busy on each (1 GiB through MemoryStream,
// tools/e2e_cpp: 64 + 63 ms with six chunks, 86 + 78 ms with sixteen).
Chunks plan_chunks(const b2rc_ctx* ctx, u64 n, u32 block, bool coarse = false)
{
    Chunks ch;
    ch.nb = b2rc_nblocks(n, block);
    u64 count = n / ctx->min_chunk;
    const u64 most = (coarse && ctx->max_chunks > 6) ? 6 : ctx->max_chunks;
    if(count > most) {
        count = most;
    }
    if(count < 1) {
        count = 1;
    }
    const u64 units = (ch.nb + 31) / 32;  // whole warps of blocks
    if(count > units) {
        count = units;
    }
    static const u32 ramp[3] = {35, 50, 70};  // per cent of a middle chunk
    u32 w[B2RC_PIPE_CHUNKS];
    u64 wsum = 0;
    for(u64 c = 0; c < count; ++c) {
        const u64 edge = c < count - 1 - c ? c : count - 1 - c;
        w[c] = (count >= 8 && edge < 3 && ctx->ramp_chunks) ? ramp[edge] : 100u;
        wsum += w[c];
    }
    u64 acc = 0, at = 0;
    ch.bnd[0] = 0;
    ch.count = 0;
    for(u64 c = 0; c < count; ++c) {
        acc += w[c];
        u64 end = c + 1 == count ? units : (units * acc + wsum / 2) / wsum;
        if(end <= at) {
            continue;  // rounding left nothing for this one
        }
        at = end;
        ch.bnd[++ch.count] = at * 32 < ch.nb ? at * 32 : ch.nb;
    }
    return ch;
}
}  // namespace

// B2RC_TRACE=1: a timeline of one b2rc_encode / b2rc_decode on stderr (device times by events against the call's
// first event, host times by the steady clock) -- tuning aid for the pipeline, off the product path otherwise.
namespace
{
struct PipeTrace {
    bool on = false;
    cudaEvent_t t0 = nullptr, a[B2RC_PIPE_CHUNKS] = {}, b[B2RC_PIPE_CHUNKS] = {}, c[B2RC_PIPE_CHUNKS] = {};
    std::chrono::steady_clock::time_point h0;
    double host[B2RC_PIPE_CHUNKS + 3] = {};
    explicit PipeTrace(cudaStream_t s0)
    {
        const char* e = getenv("B2RC_TRACE");
        on = e && atol(e) != 0;
        if(on) {
            cudaEventCreate(&t0);
            for(int k = 0; k < B2RC_PIPE_CHUNKS; ++k) {
                cudaEventCreate(&a[k]);
                cudaEventCreate(&b[k]);
                cudaEventCreate(&c[k]);
            }
            h0 = std::chrono::steady_clock::now();
            cudaEventRecord(t0, s0);
        }
    }
    void rec(cudaEvent_t* ev, u64 k, cudaStream_t st)
    {
        if(on) {
            cudaEventRecord(ev[k], st);
        }
    }
    void mark(u64 k)
    {
        if(on) {
            host[k] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
        }
    }
    void print(const char* what, u64 count, const char* la, const char* lb, const char* lc)
    {
        if(!on) {
            return;
        }
        fprintf(stderr, "[b2rc trace] %s: enqueued at %.3f ms (host), done at %.3f ms (host)\n", what, host[B2RC_PIPE_CHUNKS],
                host[B2RC_PIPE_CHUNKS + 1]);
        for(u64 k = 0; k < count; ++k) {
            float x = -1, y = -1, z = -1;
            cudaEventElapsedTime(&x, t0, a[k]);
            cudaEventElapsedTime(&y, t0, b[k]);
            cudaEventElapsedTime(&z, t0, c[k]);
            fprintf(stderr, "[b2rc trace]   chunk %2llu: %s %.3f  %s %.3f  %s %.3f  host saw it at %.3f\n", (unsigned long long)k,
                    la, x, lb, y, lc, z, host[k]);
        }
    }
    ~PipeTrace()
    {
        if(on) {
            cudaEventDestroy(t0);
            for(int k = 0; k < B2RC_PIPE_CHUNKS; ++k) {
                cudaEventDestroy(a[k]);
                cudaEventDestroy(b[k]);
                cudaEventDestroy(c[k]);
            }
        }
    }
};
}  // namespace

static int multi_encode(b2rc_ctx* ctx, int mode, u32 block_size, const u8* src, u64 n, u8* dst, u64 dst_cap, u64* out_n);
static int multi_decode(b2rc_ctx* ctx, const u8* src, u64 n, u8* dst, u64 dst_cap, u64* out_n);

int b2rc_encode(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* src, uint64_t n, uint8_t* dst,
                uint64_t dst_cap, uint64_t* out_n)
{
    if(!ctx || !dst || (n && !src) || !mode_ok(mode) || !block_ok(block_size)) {
        return B2RC_E_ARG;
    }
    if(ctx->ndev > 1 && !ctx->in_multi && b2rc_nblocks(n, block_size) >= 2ull * (u64)ctx->ndev) {
        ctx->in_multi = 1;
        const int rc = multi_encode(ctx, mode, block_size, src, n, dst, dst_cap, out_n);
        ctx->in_multi = 0;
        return rc;
    }
    DeviceGuard g(ctx->device);
    const u64 bound = b2rc_bound(mode, n, block_size);
    const bool stage_src = n >= (16ull << 20) && is_pageable(src);
    const Chunks ch = plan_chunks(ctx, n, block_size, stage_src);
    const u64 nb = ch.nb;
    const u64 idx = index_bytes(nb);
    if(out_n) {
        *out_n = bound;
    }
    if(dst_cap < idx) {
        return B2RC_E_DST_SMALL;
    }
    const u64 stride = b2rc_slot_bytes_for(mode, block_size);
    const bool need_hist = mode == B2RC_MODE_STATIC;
    const u32 seg = nb ? seg_for(ctx, mode, block_size, nb) : 0u;
    const u32 nrec = seg ? b2rc_restart_records(block_size, seg) : 0u;
    const u32 P = (mode == B2RC_MODE_STATIC && nb) ? enc_seg_plan(ctx, block_size, seg) : 0u;  // b2rc_encseg.cuh
    int rc;
    const u64 rw = rec_words(mode, block_size);
    if(seg && (rc = grow(ctx, ctx->restart, ctx->restart_cap, (size_t)(nb * nrec * 4ull * rw + 16))) != B2RC_OK) {
        return rc;
    }
    if((rc = grow(ctx, ctx->stage_in, ctx->stage_in_cap, (size_t)(n + 16))) != B2RC_OK ||
       (rc = grow(ctx, ctx->stage_out, ctx->stage_out_cap, (size_t)(bound + 16))) != B2RC_OK ||
       (!P && (rc = grow(ctx, ctx->slots, ctx->slots_cap, (size_t)(nb * stride))) != B2RC_OK) ||
       (P && (rc = seg_scratch(ctx, nb, block_size, P)) != B2RC_OK) ||
       (rc = grow(ctx, ctx->sizes, ctx->sizes_cap, (size_t)(nb * 4 + 16))) != B2RC_OK ||
       (need_hist && (rc = grow(ctx, ctx->freq16, ctx->freq_cap, (size_t)(nb * 512 + 16))) != B2RC_OK)) {
        return rc;
    }
    u8* d_payload = ctx->stage_out + idx;
    u64* d_offsets = reinterpret_cast<u64*>(ctx->stage_out + B2RC_HEADER_BYTES);
    cudaStream_t s0 = ctx->pipe[0];
    if(stage_src && (rc = grow_host(ctx, ctx->h_in, ctx->h_in_cap, (size_t)n + 16)) != B2RC_OK) {
        return rc;
    }
    CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), s0));
    CK(cudaMemsetAsync(ctx->d_ends, 0, 8, s0));
    if(nb == 0) {
        CK(cudaMemsetAsync(d_offsets, 0, 8, s0));
    }
    CK(cudaEventRecord(ctx->index_ready, s0));
    PipeTrace tr(s0);
    // enqueue every chunk
    for(u64 c = 0; c < ch.count; ++c) {
        cudaStream_t st = ctx->pipe[c % B2RC_PIPE_STREAMS];
        const u64 b0 = ch.lo(c), b1 = ch.hi(c);
        const u64 byte0 = b0 * block_size;
        const u64 bytes = (b1 * (u64)block_size < n ? b1 * (u64)block_size : n) - byte0;
        if(c == 0 || c < B2RC_PIPE_STREAMS) {
            CK(cudaStreamWaitEvent(st, ctx->index_ready, 0));
        }
        if(stage_src) {
            b2rc_host_copy(ctx->h_in + byte0, src + byte0, bytes);  // chunk c-1 is on the wire meanwhile
        }
        CK(cudaMemcpyAsync(ctx->stage_in + byte0, (stage_src ? ctx->h_in : src) + byte0, bytes, cudaMemcpyHostToDevice, st));
        tr.rec(tr.a, c, st);
        u16* freq = need_hist ? ctx->freq16 + b0 * 256 : nullptr;
        if(need_hist && (rc = b2rc_k_histogram(ctx, ctx->stage_in + byte0, bytes, block_size, freq, st)) != B2RC_OK) {
            return rc;
        }
        SegArgs sa;
        if(P) {  // sizes first (the range pass), the payloads then go straight to their final place
            sa = seg_args(ctx, block_size, P, ctx->stage_in + byte0, bytes, freq, b0, ctx->sizes + b0, d_offsets + b0,
                          d_payload, bound - idx, seg ? ctx->restart + b0 * nrec * rw : nullptr, seg);
            // the chunks arrive as fast as the wire brings them: one or two are being coded at any time, the GPU is
            // mostly idle, and the range pass may take its helper warps (its latency is the pipeline's drain)
            if((rc = static_ranges_launch(ctx, sa, 2 * (b1 - b0), st)) != B2RC_OK) {
                return rc;
            }
        } else if((rc = b2rc_k_encode_blocks_r(ctx, mode, block_size, ctx->stage_in + byte0, bytes, freq,
                                               ctx->slots + b0 * stride, stride, ctx->sizes + b0,
                                               seg ? ctx->restart + b0 * nrec * rw : nullptr, seg, ctx->d_err, st)) !=
                  B2RC_OK) {
            return rc;
        }
        if(c > 0) {
            CK(cudaStreamWaitEvent(st, ctx->scan_done[c - 1], 0));
        }
        if((rc = scan_launch(ctx, ctx->sizes + b0, b1 - b0, d_offsets + b0, ctx->d_ends + c + 1, nullptr, 0, 0, 0, st,
                             ctx->d_ends + c)) != B2RC_OK) {
            return rc;
        }
        CK(cudaEventRecord(ctx->scan_done[c], st));
        if(P) {
            rc = static_segments_launch(ctx, sa, st);
        } else {
            rc = b2rc_k_compact_for(ctx, mode, ctx->slots + b0 * stride, stride, ctx->sizes + b0, d_offsets + b0, b1 - b0,
                                    d_payload, bound - idx, ctx->d_err, st);
        }
        if(rc != B2RC_OK) {
            return rc;
        }
        tr.rec(tr.b, c, st);
        CK(cudaMemcpyAsync(ctx->h_ends + c + 1, ctx->d_ends + c + 1, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->chunk_done[c], st));
    }
    tr.mark(B2RC_PIPE_CHUNKS);
    // drain: as each chunk's size becomes known, send its payload home on the chunk's own stream
    ctx->h_ends[0] = 0;
    int result = B2RC_OK;
    for(u64 c = 0; c < ch.count; ++c) {
        CK(cudaEventSynchronize(ctx->chunk_done[c]));
        tr.mark(c);
        const u64 lo = ctx->h_ends[c], hi = ctx->h_ends[c + 1];
        if(idx + hi > dst_cap) {
            result = B2RC_E_DST_SMALL;
            continue;
        }
        if(result == B2RC_OK && hi > lo) {
            // on the context's own stream: the chunk's stream still has later chunks queued
            CK(cudaMemcpyAsync(dst + idx + lo, d_payload + lo, hi - lo, cudaMemcpyDeviceToHost, ctx->stream));
        }
        tr.rec(tr.c, c, ctx->stream);
    }
    const u64 total = ch.count ? ctx->h_ends[ch.count] : 0;
    const u64 table_bytes = (u64)nb * nrec * 4ull * rw;
    const u64 made = idx + (seg ? align4(total) + table_bytes : total);
    if(out_n) {
        *out_n = made;
    }
    if(seg && result == B2RC_OK) {
        if(made > dst_cap) {
            result = B2RC_E_DST_SMALL;
        } else {
            // every chunk_done event has been waited for: the records are complete
            memset(dst + idx + total, 0, (size_t)(align4(total) - total));
            CK(cudaMemcpyAsync(dst + idx + align4(total), ctx->restart, table_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    // header from the host, index from the device
    u32 h[8] = {0x43523242u, 1u | ((u32)mode << 16), block_size, flags_of(seg), (u32)n, (u32)(n >> 32), (u32)nb,
                (u32)(nb >> 32)};
    memcpy(dst, h, sizeof h);
    CK(cudaMemcpyAsync(&ctx->h_res->err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, s0));
    if(result == B2RC_OK) {
        CK(cudaMemcpyAsync(dst + B2RC_HEADER_BYTES, d_offsets, 8 * (nb + 1), cudaMemcpyDeviceToHost, s0));
    }
    for(int k = 0; k < B2RC_PIPE_STREAMS; ++k) {
        CK(cudaStreamSynchronize(ctx->pipe[k]));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    tr.mark(B2RC_PIPE_CHUNKS + 1);
    tr.print("b2rc_encode", ch.count, "in", "coded", "home");
    const int kerr = map_kernel_err(ctx->h_res->err);
    return kerr != B2RC_OK ? kerr : result;
}

int b2rc_decode(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* out_n)
{
    if(!ctx || !src) {
        return B2RC_E_ARG;
    }
    int mode;
    u32 block;
    u64 total, nb;
    int rc = b2rc_peek(src, n, &mode, &block, &total, &nb);
    if(rc != B2RC_OK) {
        return rc;
    }
    if(ctx->ndev > 1 && !ctx->in_multi && nb >= 2ull * (u64)ctx->ndev) {
        ctx->in_multi = 1;
        rc = multi_decode(ctx, src, n, dst, dst_cap, out_n);
        ctx->in_multi = 0;
        return rc;
    }
    if(out_n) {
        *out_n = total;
    }
    if(dst_cap < total || (total && !dst)) {
        return B2RC_E_DST_SMALL;
    }
    // the index is in host memory here: check it (b2rc_check) before anything is allocated or reaches the device
    const u64 idx = index_bytes(nb);
    const u64 payload_len = n - idx;
    if((rc = b2rc_check(src, n, nullptr)) != B2RC_OK) {
        return rc;
    }
    u64 prev = 0;
    memcpy(&prev, src + B2RC_HEADER_BYTES + 8 * nb, 8);
    if(nb == 0) {
        return B2RC_OK;
    }
    // a restart table behind the payloads (static coder): the decoder then runs a chain per segment
    u32 flags;
    memcpy(&flags, src + 12, 4);
    const u32 seg = flags ? (flags >> 8) * 64u : 0u;
    const u32 nrec = seg ? b2rc_restart_records(block, seg) : 0u;
    const u64 rw = rec_words(mode, block);
    const u64 table_bytes = (u64)nb * nrec * 4ull * rw;
    if(seg && (align4(prev) + table_bytes > payload_len)) {
        return B2RC_E_CORRUPT;
    }
    DeviceGuard g(ctx->device);
    if((rc = grow(ctx, ctx->stage_in, ctx->stage_in_cap, (size_t)(n + 16))) != B2RC_OK ||
       (rc = grow(ctx, ctx->stage_out, ctx->stage_out_cap, (size_t)(total + 16))) != B2RC_OK ||
       (seg && (rc = grow(ctx, ctx->restart, ctx->restart_cap, (size_t)(table_bytes + 16))) != B2RC_OK)) {
        return rc;
    }
    const bool stage_src = n >= (16ull << 20) && is_pageable(src);
    const Chunks ch = plan_chunks(ctx, total, block, stage_src);
    if(stage_src && (rc = grow_host(ctx, ctx->h_in, ctx->h_in_cap, (size_t)n + 16)) != B2RC_OK) {
        return rc;
    }
    // range coders, long blocks, more than one chunk in flight: decode in phases (see below)
    u32 phases = 1, per = block;
    const u64 model_bytes = 512ull * 32ull * (block > 65536u ? 4u : 2u);  // per warp of 32 blocks
    if(!is_ans(mode) && !seg && block >= 2u * B2RC_PHASE_MIN_SYMS && ch.count > 1 && ctx->max_phases > 1) {
        phases = block / B2RC_PHASE_MIN_SYMS;
        if(phases > ctx->max_phases) {
            phases = (u32)ctx->max_phases;
        }
        per = ((block / phases) + 63u) & ~63u;
        phases = (block + per - 1) / per;
        if((rc = grow(ctx, ctx->dec_state, ctx->dec_state_cap, (size_t)(nb * 32 + 64))) != B2RC_OK ||
           (mode == B2RC_MODE_ADAPTIVE &&
            (rc = grow(ctx, ctx->dec_model, ctx->dec_model_cap, (size_t)(((nb + 31) / 32) * model_bytes))) != B2RC_OK)) {
            return rc;
        }
    }
    cudaStream_t s0 = ctx->pipe[0];
    const u64* d_offsets = reinterpret_cast<const u64*>(ctx->stage_in + B2RC_HEADER_BYTES);
    CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), s0));
    CK(cudaMemcpyAsync(ctx->stage_in, src, idx, cudaMemcpyHostToDevice, s0));  // header + index
    if(seg) {
        CK(cudaMemcpyAsync(ctx->restart, src + idx + align4(prev), table_bytes, cudaMemcpyHostToDevice, s0));
    }
    CK(cudaEventRecord(ctx->index_ready, s0));
    PipeTrace tr(s0);
    for(u64 c = 0; c < ch.count; ++c) {
        cudaStream_t st = ctx->pipe[c % B2RC_PIPE_STREAMS];
        const u64 b0 = ch.lo(c), b1 = ch.hi(c);
        u64 p0, p1;
        memcpy(&p0, src + B2RC_HEADER_BYTES + 8 * b0, 8);
        memcpy(&p1, src + B2RC_HEADER_BYTES + 8 * b1, 8);
        const u64 byte0 = b0 * (u64)block;
        const u64 bytes = (b1 * (u64)block < total ? b1 * (u64)block : total) - byte0;
        if(c < B2RC_PIPE_STREAMS) {
            CK(cudaStreamWaitEvent(st, ctx->index_ready, 0));
        }
        if(p1 > p0) {
            if(stage_src) {
                b2rc_host_copy(ctx->h_in + idx + p0, src + idx + p0, p1 - p0);
            }
            CK(cudaMemcpyAsync(ctx->stage_in + idx + p0, (stage_src ? ctx->h_in : src) + idx + p0, p1 - p0,
                               cudaMemcpyHostToDevice, st));
        }
        tr.rec(tr.a, c, st);
        if(phases <= 1) {
            if((rc = b2rc_k_decode_blocks_r(ctx, mode, block, ctx->stage_in + idx, payload_len, d_offsets + b0, b1 - b0,
                                            ctx->stage_out + byte0, bytes, seg ? ctx->restart + b0 * nrec * rw : nullptr,
                                            seg, ctx->d_err, st)) != B2RC_OK) {
                return rc;
            }
            tr.rec(tr.b, c, st);
            CK(cudaMemcpyAsync(dst + byte0, ctx->stage_out + byte0, bytes, cudaMemcpyDeviceToHost, st));
            tr.rec(tr.c, c, st);
            continue;
        }
        // Phased: a launch decodes the next `per` symbols of every block of the chunk, and that
        // stripe of the output goes home (a 2-D copy, one row per block) while the next launch
        // runs.  A static decode launch takes as long for one block as for a thousand -- the
        // chain is serial per block -- so without this the first byte of a chunk's output would
        // only leave the GPU when the whole chunk is done.
        const u64 full_rows = bytes / block;          // blocks of the chunk that are complete
        const u32 tail = (u32)(bytes - full_rows * block);  // the stream's ragged last block, if here
        for(u32 p = 0; p < phases; ++p) {
            const u32 s0 = p * per;
            const u32 ns = (p + 1 == phases) ? (((block - s0) + 63u) & ~63u) : per;
            if((rc = decode_launch(ctx, mode, block, ctx->stage_in + idx, payload_len, d_offsets + b0, b1 - b0,
                                   ctx->stage_out + byte0, bytes, ctx->d_err, st, s0, ns, ctx->dec_state + b0 * 8,
                                   mode == B2RC_MODE_ADAPTIVE ? ctx->dec_model + (b0 / 32) * model_bytes : nullptr)) !=
               B2RC_OK) {
                return rc;
            }
            CK(cudaEventRecord(ctx->phase_done[c][p], st));
            cudaStream_t out = ctx->d2h[p];
            CK(cudaStreamWaitEvent(out, ctx->phase_done[c][p], 0));
            const u32 width = (s0 + ns <= block) ? ns : block - s0;
            if(full_rows) {
                CK(cudaMemcpy2DAsync(dst + byte0 + s0, block, ctx->stage_out + byte0 + s0, block, width, full_rows,
                                     cudaMemcpyDeviceToHost, out));
            }
            if(tail > s0) {
                const u32 tw = tail - s0 < width ? tail - s0 : width;
                CK(cudaMemcpyAsync(dst + byte0 + full_rows * block + s0, ctx->stage_out + byte0 + full_rows * block + s0,
                                   tw, cudaMemcpyDeviceToHost, out));
            }
        }
    }
    for(int k = 1; k < B2RC_PIPE_STREAMS; ++k) {
        CK(cudaStreamSynchronize(ctx->pipe[k]));
    }
    for(int k = 0; k < B2RC_PHASES; ++k) {
        CK(cudaStreamSynchronize(ctx->d2h[k]));
    }
    CK(cudaMemcpyAsync(&ctx->h_res->err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    tr.mark(B2RC_PIPE_CHUNKS + 1);
    if(phases <= 1) {
        tr.print("b2rc_decode", ch.count, "in", "decoded", "home");
    }
    return map_kernel_err(ctx->h_res->err);
}

// ------------------------------------------------------------ several devices --
// b2rc_ctx_create_multi: the host-pointer calls over the devices of one process.  Device d takes the
// blocks [d * nb / N, (d + 1) * nb / N) on a host thread of its own and codes them with
// the single-device call into its pinned staging; the sizes meet on the host (no collective), the
// pieces are stitched by a few copy threads into the container a single device would have written.
namespace
{
struct Piece {
    u64 blk_lo, blk_hi;
    const u8* out;   // a container of its own in the device's pinned staging
    u64 made;
    int rc;
};
u64 rd64(const u8* p)
{
    u64 v;
    memcpy(&v, p, 8);
    return v;
}
}  // namespace

static int multi_encode(b2rc_ctx* ctx, int mode, u32 block_size, const u8* src, u64 n, u8* dst, u64 dst_cap, u64* out_n)
{
    const int N = ctx->ndev;
    const u64 nb = b2rc_nblocks(n, block_size);
    const u64 idx = index_bytes(nb);
    if(out_n) {
        *out_n = b2rc_bound(mode, n, block_size);
    }
    if(dst_cap < idx) {
        return B2RC_E_DST_SMALL;
    }
    // the spacing of the restart points follows the share of ONE device (each decodes its share on its own)
    const u32 seg_each = ctx->seg_force ? ctx->seg_force : seg_for(ctx, mode, block_size, (nb + (u64)N - 1) / (u64)N);
    std::vector<Piece> pc((size_t)N);
    std::vector<std::thread> pool;
    for(int d = 0; d < N; ++d) {
        pc[d].blk_lo = nb * (u64)d / (u64)N;
        pc[d].blk_hi = nb * (u64)(d + 1) / (u64)N;
        pool.emplace_back([&, d] {
            const u64 lo = pc[d].blk_lo * block_size;
            const u64 hi = pc[d].blk_hi * (u64)block_size < n ? pc[d].blk_hi * (u64)block_size : n;
            const u32 keep = ctx->sub[d]->seg_force;  // sub[0] is the head itself
            ctx->sub[d]->seg_force = seg_each;        // every piece with the same spacing: the pieces are stitched
            pc[d].rc = b2rc_encode_staged(ctx->sub[d], mode, block_size, src + lo, hi - lo, &pc[d].out, &pc[d].made);
            ctx->sub[d]->seg_force = keep;
        });
    }
    for(auto& th : pool) {
        th.join();
    }
    u64 total = 0, table = 0;
    u32 flags = 0;
    for(int d = 0; d < N; ++d) {
        if(pc[d].rc != B2RC_OK) {
            snprintf(ctx->last_err, sizeof ctx->last_err, "device %d: %.200s", ctx->sub[d]->device, ctx->sub[d]->last_err);
            return pc[d].rc;
        }
        const u64 nbd = pc[d].blk_hi - pc[d].blk_lo;
        const u64 pay = rd64(pc[d].out + B2RC_HEADER_BYTES + 8 * nbd);
        memcpy(&flags, pc[d].out + 12, 4);
        total += pay;
        table += pc[d].made - index_bytes(nbd) - (flags ? align4(pay) : pay);
    }
    const u64 made = idx + (flags ? align4(total) + table : total);
    if(out_n) {
        *out_n = made;
    }
    if(made > dst_cap) {
        return B2RC_E_DST_SMALL;
    }
    const u32 h[8] = {0x43523242u, 1u | ((u32)mode << 16), block_size, flags, (u32)n, (u32)(n >> 32), (u32)nb, (u32)(nb >> 32)};
    memcpy(dst, h, sizeof h);
    memset(dst + idx + total, 0, (size_t)((flags ? align4(total) : total) - total));
    pool.clear();
    u64 base = 0, tab_at = idx + align4(total);
    for(int d = 0; d < N; ++d) {
        const u64 nbd = pc[d].blk_hi - pc[d].blk_lo;
        const u8* index = pc[d].out + B2RC_HEADER_BYTES;
        const u64 pay = rd64(index + 8 * nbd);
        const u64 tab = pc[d].made - index_bytes(nbd) - (flags ? align4(pay) : pay);
        const u8* from = pc[d].out;
        const u64 lo = pc[d].blk_lo;
        pool.emplace_back([=] {
            for(u64 b = 0; b <= nbd; ++b) {  // the last entry of a piece is the first of the next: the same value
                const u64 o = rd64(index + 8 * b) + base;
                memcpy(dst + B2RC_HEADER_BYTES + 8 * (lo + b), &o, 8);
            }
            b2rc_host_copy(dst + idx + base, from + index_bytes(nbd), pay);
            if(tab) {
                memcpy(dst + tab_at, from + index_bytes(nbd) + align4(pay), (size_t)tab);
            }
        });
        base += pay;
        tab_at += tab;
    }
    for(auto& th : pool) {
        th.join();
    }
    return B2RC_OK;
}

static int multi_decode(b2rc_ctx* ctx, const u8* src, u64 n, u8* dst, u64 dst_cap, u64* out_n)
{
    int mode;
    u32 block;
    u64 total, nb;
    int rc = b2rc_check(src, n, &total);  // header and the whole index, before anything is sized by them
    if(rc != B2RC_OK || (rc = b2rc_peek(src, n, &mode, &block, nullptr, &nb)) != B2RC_OK) {
        return rc;
    }
    if(out_n) {
        *out_n = total;
    }
    if(dst_cap < total || (total && !dst)) {
        return B2RC_E_DST_SMALL;
    }
    const int N = ctx->ndev;
    const u64 idx = index_bytes(nb);
    u32 flags;
    memcpy(&flags, src + 12, 4);
    const u32 seg = flags ? (flags >> 8) * 64u : 0u;
    const u64 rec_bytes = seg ? 4ull * rec_words(mode, block) * b2rc_restart_records(block, seg) : 0ull;  // per block
    const u64 pay_end = rd64(src + B2RC_HEADER_BYTES + 8 * nb);
    std::vector<int> rcs((size_t)N, B2RC_OK);
    std::vector<std::thread> pool;
    for(int d = 0; d < N; ++d) {
        pool.emplace_back([&, d] {
            // a container of this device's blocks alone, in its pinned staging: header, index from zero, payloads, points
            b2rc_ctx* c = ctx->sub[d];
            const u64 lo = nb * (u64)d / (u64)N, hi = nb * (u64)(d + 1) / (u64)N, nbd = hi - lo;
            const u64 p0 = rd64(src + B2RC_HEADER_BYTES + 8 * lo), p1 = rd64(src + B2RC_HEADER_BYTES + 8 * hi);
            const u64 byte_lo = lo * block, byte_hi = hi * (u64)block < total ? hi * (u64)block : total;
            const u64 need = index_bytes(nbd) + align4(p1 - p0) + rec_bytes * nbd;
            DeviceGuard g(c->device);
            if((rcs[d] = grow_host(c, c->h_in, c->h_in_cap, (size_t)need + 16)) != B2RC_OK) {
                return;
            }
            u8* sub = c->h_in;
            const u64 tot_d = byte_hi - byte_lo;
            const u32 h[8] = {0x43523242u, 1u | ((u32)mode << 16), block, flags, (u32)tot_d, (u32)(tot_d >> 32), (u32)nbd, (u32)(nbd >> 32)};
            memcpy(sub, h, sizeof h);
            for(u64 b = 0; b <= nbd; ++b) {
                const u64 o = rd64(src + B2RC_HEADER_BYTES + 8 * (lo + b)) - p0;
                memcpy(sub + B2RC_HEADER_BYTES + 8 * b, &o, 8);
            }
            b2rc_host_copy(sub + index_bytes(nbd), src + idx + p0, p1 - p0);
            if(seg) {
                memset(sub + index_bytes(nbd) + (p1 - p0), 0, (size_t)(align4(p1 - p0) - (p1 - p0)));
                memcpy(sub + index_bytes(nbd) + align4(p1 - p0), src + idx + align4(pay_end) + rec_bytes * lo, (size_t)(rec_bytes * nbd));
            }
            u64 got = 0;
            rcs[d] = b2rc_decode(c, sub, index_bytes(nbd) + (seg ? align4(p1 - p0) + rec_bytes * nbd : p1 - p0), dst + byte_lo,
                                 tot_d, &got);
        });
    }
    for(auto& th : pool) {
        th.join();
    }
    for(int d = 0; d < N; ++d) {
        if(rcs[d] != B2RC_OK) {
            snprintf(ctx->last_err, sizeof ctx->last_err, "device %d: %.200s", ctx->sub[d]->device, ctx->sub[d]->last_err);
            return rcs[d];
        }
    }
    return B2RC_OK;
}

int b2rc_container_bytes(const uint8_t* prefix, uint64_t have, uint64_t* need)
{
    if(!prefix || !need) {
        return B2RC_E_ARG;
    }
    if(have < B2RC_HEADER_BYTES) {
        *need = B2RC_HEADER_BYTES + 8;
        return B2RC_OK;
    }
    u32 h[8];
    memcpy(h, prefix, sizeof h);
    const u32 md = h[1] >> 16;
    if(h[0] != 0x43523242u || (h[1] & 0xFFFFu) != 1u || !mode_ok((int)md) || !block_ok(h[2])) {
        return B2RC_E_CORRUPT;
    }
    const u64 nb = (u64)h[6] | ((u64)h[7] << 32);
    if(nb > (1ull << 40)) {
        return B2RC_E_CORRUPT;
    }
    const u64 idx = index_bytes(nb);
    if(have < idx) {
        *need = idx;
        return B2RC_OK;
    }
    const u64 last = rd64(prefix + idx - 8);
    u64 total = idx + last;
    if(h[3] != 0u) {
        const u32 seg = (h[3] >> 8) * 64u;
        if((h[3] & 0xFFu) != 1u || !has_restart((int)md) || !seg_ok(h[2], seg)) {
            return B2RC_E_CORRUPT;
        }
        total = idx + align4(last) + nb * 4ull * rec_words((int)md, h[2]) * b2rc_restart_records(h[2], seg);
    }
    *need = total;
    return B2RC_OK;
}

// ------------------------------------------------------- host-side validation --
int b2rc_check(const uint8_t* src, uint64_t n, uint64_t* total_out)
{
    int mode;
    u32 block;
    u64 total, nb;
    const int rc = b2rc_peek(src, n, &mode, &block, &total, &nb);
    if(rc != B2RC_OK) {
        return rc;
    }
    const u64 idx = index_bytes(nb);
    const u64 payload_len = n - idx;
    const u64 min_pay = is_ans(mode) ? (u64)ANS_HDR + 4u : (mode == B2RC_MODE_STATIC ? (u64)RC_STATIC_HDR + 5u : (u64)RC_ADAPT_HDR + 5u);
    u64 prev = 0;
    for(u64 b = 0; b <= nb; ++b) {
        u64 o;
        memcpy(&o, src + B2RC_HEADER_BYTES + 8 * b, 8);
        if(o < prev || o > payload_len || (b == 0 && o != 0) || (b > 0 && o - prev < min_pay)) {
            return B2RC_E_CORRUPT;
        }
        prev = o;
    }
    u32 flags;
    memcpy(&flags, src + 12, 4);
    if(flags) {
        const u32 seg = (flags >> 8) * 64u;
        if(align4(prev) + (u64)nb * b2rc_restart_records(block, seg) * 4ull * rec_words(mode, block) > payload_len) {
            return B2RC_E_CORRUPT;
        }
    }
    if(total_out) {
        *total_out = total;
    }
    return B2RC_OK;
}

static int grow_host(b2rc_ctx* ctx, u8*& buf, size_t& cap, size_t want)
{
    if(want <= cap) {
        return B2RC_OK;
    }
    if(buf) {
        CK(cudaFreeHost(buf));
        buf = nullptr;
        cap = 0;
    }
    void* q = nullptr;
    if(!cuda_ok(ctx, cudaMallocHost(&q, want), "cudaMallocHost")) {
        return B2RC_E_NOMEM;
    }
    buf = static_cast<u8*>(q);
    cap = want;
    return B2RC_OK;
}

int b2rc_encode_staged(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* src, uint64_t n,
                       const uint8_t** out, uint64_t* out_n)
{
    if(!ctx || !out || !out_n || !mode_ok(mode) || !block_ok(block_size)) {
        return B2RC_E_ARG;
    }
    DeviceGuard g(ctx->device);
    const int rc = grow_host(ctx, ctx->h_stage, ctx->h_stage_cap, (size_t)b2rc_bound(mode, n, block_size) + 16);
    if(rc != B2RC_OK) {
        return rc;
    }
    *out = ctx->h_stage;
    return b2rc_encode(ctx, mode, block_size, src, n, ctx->h_stage, ctx->h_stage_cap, out_n);
}

int b2rc_decode_staged(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, const uint8_t** out, uint64_t* out_n)
{
    if(!ctx || !src || !out || !out_n) {
        return B2RC_E_ARG;
    }
    u64 total = 0;
    int rc = b2rc_check(src, n, &total);  // before any allocation sized by the header
    if(rc != B2RC_OK) {
        return rc;
    }
    DeviceGuard g(ctx->device);
    if((rc = grow_host(ctx, ctx->h_stage, ctx->h_stage_cap, (size_t)total + 16)) != B2RC_OK) {
        return rc;
    }
    *out = ctx->h_stage;
    return b2rc_decode(ctx, src, n, ctx->h_stage, ctx->h_stage_cap, out_n);
}

void b2rc_host_copy(void* dst, const void* src, uint64_t n)
{
    // a single thread copies at a third of the PCIe rate; the caller waits for this copy and nothing else runs on its
    // behalf meanwhile, so every hardware thread up to 16 takes a share (1 GiB through MemoryStream, tools/e2e_cpp:
    // 4 / 8 / 16 threads 5.3 / 6.9 / 8.4 GB/s)
    const unsigned hw = std::thread::hardware_concurrency();
    unsigned threads = hw >= 16 ? 16u : (hw >= 8 ? hw : (hw >= 4 ? 4u : (hw >= 2 ? 2u : 1u)));
    if(const char* e = getenv("B2RC_COPY_THREADS")) {  // tuning
        const long v = atol(e);
        if(v >= 1 && v <= 64) {
            threads = (unsigned)v;
        }
    }
    if(n < (32ull << 20)) {
        threads = 1;
    }
    if(threads == 1) {
        memcpy(dst, src, (size_t)n);
        return;
    }
    const u64 per = ((n / threads) + 4095ull) & ~4095ull;
    std::vector<std::thread> pool;
    for(unsigned t = 1; t < threads; ++t) {
        const u64 lo = per * t, hi = (t + 1 == threads) ? n : (per * (t + 1) < n ? per * (t + 1) : n);
        if(lo < hi) {
            pool.emplace_back([=] { memcpy((u8*)dst + lo, (const u8*)src + lo, (size_t)(hi - lo)); });
        }
    }
    memcpy(dst, src, (size_t)(per < n ? per : n));
    for(auto& th : pool) {
        th.join();
    }
}

int b2rc_host_alloc(uint64_t bytes, void** out)
{
    if(!out) {
        return B2RC_E_ARG;
    }
    *out = nullptr;
    void* q = nullptr;
    if(cudaMallocHost(&q, bytes ? (size_t)bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        return B2RC_E_NOMEM;
    }
    *out = q;
    return B2RC_OK;
}

void b2rc_host_free(void* p)
{
    if(p) {
        cudaFreeHost(p);
        cudaGetLastError();
    }
}

// ------------------------------------------------------------- block sort --
// blksort::BlkSort behind the C ABI (include/b2rc.h, "block sort").  Sizes are a function of n:
// full 32 KiB blocks grow by two bytes, the rest is copied.
uint64_t b2rc_blk_encode_bound(uint64_t n)
{
    const u64 blocks = n >> 15;
    return blocks * BLK_CODED + (n - (blocks << 15));
}

uint64_t b2rc_blk_decoded_size(uint64_t n)
{
    const u64 blocks = n / BLK_CODED;
    return blocks * BLK_N + (n - blocks * BLK_CODED);
}

static int blk_forward_launch(b2rc_ctx* ctx, const u8* d_src, u8* d_dst, u64 b0, u64 nb, cudaStream_t st)
{
    u64 done = 0;
    while(done < nb) {  // a grid dimension holds 2^31 - 1 blocks: 64 TiB per launch, the loop is for form
        const u64 now = nb - done < 0x7FFFFFFFull ? nb - done : 0x7FFFFFFFull;
        KernelTimer kt(ctx, B2RC_K_BLK_FORWARD, st);
        k_blk_fwd<<<(unsigned)now, BLK_THREADS, BLK_FWD_SMEM, st>>>(d_src + (b0 + done) * BLK_N, d_dst + (b0 + done) * BLK_CODED,
                                                                    ctx->blk_rounds + b0 + done, nullptr, nullptr,
                                                                    ctx->blk_tie_list, (u32)(b0 + done));
        const int rc = launch_check(ctx, "k_blk_fwd");
        if(rc != B2RC_OK) {
            return rc;
        }
        done += now;
    }
    return B2RC_OK;
}

// Blocks with a period: the row number B1 wrote is the canonical one; the reference's depends on the swaps of
// its quicksort.  k_blk_ties replays them (b2rc_blk.cuh, B3).  B1 appended such blocks to a device list through
// an atomic counter: the host reads the counter alone (one word, the stream drains for it) and does nothing more
// when it is zero -- no copy of per-block words, no host work proportional to the number of blocks.
// `fixed` (optional) receives the blocks whose row number was rewritten.
static int blk_fix_ties(b2rc_ctx* ctx, const u8* d_src, u8* d_dst, u64 nb, cudaStream_t st, std::vector<u64>* fixed)
{
    if(nb == 0) {
        return B2RC_OK;
    }
    CK(cudaMemcpyAsync(&ctx->h_res->pad, ctx->blk_tie_list, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const u32 count = (u32)ctx->h_res->pad;
    if(count == 0) {
        return B2RC_OK;
    }
    std::vector<u32> list((size_t)count);
    CK(cudaMemcpyAsync(list.data(), ctx->blk_tie_list + 1, (size_t)count * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::sort(list.begin(), list.end());  // the counter hands places out in arrival order
    const size_t batch = list.size() < 592 ? list.size() : 592;  // four waves of one CTA per SM; 320 KiB of scratch per block
    const size_t rk_bytes = (size_t)BLK_N * 2, q_bytes = (size_t)2 * BLK_TIES_QUEUE * sizeof(TieRange);
    int rc;
    if((rc = grow(ctx, ctx->blk_ties, ctx->blk_ties_cap, batch * (rk_bytes + q_bytes + 4) + 256)) != B2RC_OK) {
        return rc;
    }
    u16* d_rk = reinterpret_cast<u16*>(ctx->blk_ties);
    TieRange* d_q = reinterpret_cast<TieRange*>(ctx->blk_ties + batch * rk_bytes);
    u32* d_list = reinterpret_cast<u32*>(ctx->blk_ties + batch * (rk_bytes + q_bytes));
    for(size_t at = 0; at < list.size(); at += batch) {
        const size_t now = list.size() - at < batch ? list.size() - at : batch;
        CK(cudaMemcpyAsync(d_list, list.data() + at, now * 4, cudaMemcpyHostToDevice, st));
        k_blk_fwd<<<(unsigned)now, BLK_THREADS, BLK_FWD_SMEM, st>>>(d_src, d_dst, ctx->blk_rounds, d_list, d_rk);
        if((rc = launch_check(ctx, "k_blk_fwd (ranks)")) != B2RC_OK) {
            return rc;
        }
        k_blk_ties<<<(unsigned)now, BLK_THREADS, BLK_TIES_SMEM, st>>>(d_src, d_dst, d_list, d_rk, d_q);
        if((rc = launch_check(ctx, "k_blk_ties")) != B2RC_OK) {
            return rc;
        }
        CK(cudaStreamSynchronize(st));  // d_list is reused by the next batch
    }
    if(fixed) {
        fixed->assign(list.begin(), list.end());
    }
    return B2RC_OK;
}

static int blk_inverse_launch(b2rc_ctx* ctx, const u8* d_src, u8* d_dst, u64 b0, u64 nb, cudaStream_t st)
{
    u64 done = 0;
    while(done < nb) {
        const u64 now = nb - done < 0x7FFFFFFFull ? nb - done : 0x7FFFFFFFull;
        KernelTimer kt(ctx, B2RC_K_BLK_INVERSE, st);
        k_blk_inv<<<(unsigned)now, BLK_THREADS, BLK_INV_SMEM, st>>>(d_src + (b0 + done) * BLK_CODED, d_dst + (b0 + done) * BLK_N,
                                                                    ctx->d_err);
        const int rc = launch_check(ctx, "k_blk_inv");
        if(rc != B2RC_OK) {
            return rc;
        }
        done += now;
    }
    return B2RC_OK;
}

int b2rc_blk_encode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                           uint64_t* out_n, void* cuda_stream)
{
    if(!ctx || (n && (!d_src || !d_dst)) || !aligned16(d_src) || !aligned16(d_dst)) {
        return B2RC_E_ARG;
    }
    const u64 need = b2rc_blk_encode_bound(n);
    if(out_n) {
        *out_n = need;
    }
    if(dst_cap < need) {
        return B2RC_E_DST_SMALL;
    }
    const u64 nb = n >> 15;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc;
    if((rc = grow(ctx, ctx->blk_rounds, ctx->blk_rounds_cap, (size_t)(nb * 4 + 16))) != B2RC_OK ||
       (rc = grow(ctx, ctx->blk_tie_list, ctx->blk_tie_cap, (size_t)(nb * 4 + 16))) != B2RC_OK) {
        return rc;
    }
    ctx->blk_last_blocks = nb;
    CK(cudaMemsetAsync(ctx->blk_tie_list, 0, 4, st));
    if(nb && (rc = blk_forward_launch(ctx, d_src, d_dst, 0, nb, st)) != B2RC_OK) {
        return rc;
    }
    if(n > nb * BLK_N) {
        CK(cudaMemcpyAsync(d_dst + nb * BLK_CODED, d_src + nb * BLK_N, (size_t)(n - nb * BLK_N), cudaMemcpyDeviceToDevice, st));
    }
    return blk_fix_ties(ctx, d_src, d_dst, nb, st, nullptr);  // drains the stream
}

int b2rc_blk_decode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                           uint64_t* out_n, void* cuda_stream)
{
    if(!ctx || (n && (!d_src || !d_dst)) || !aligned16(d_src) || !aligned16(d_dst)) {
        return B2RC_E_ARG;
    }
    const u64 need = b2rc_blk_decoded_size(n);
    if(out_n) {
        *out_n = need;
    }
    if(dst_cap < need) {
        return B2RC_E_DST_SMALL;
    }
    const u64 nb = n / BLK_CODED;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc;
    CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), st));
    if(nb && (rc = blk_inverse_launch(ctx, d_src, d_dst, 0, nb, st)) != B2RC_OK) {
        return rc;
    }
    if(n > nb * BLK_CODED) {
        CK(cudaMemcpyAsync(d_dst + nb * BLK_N, d_src + nb * BLK_CODED, (size_t)(n - nb * BLK_CODED), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaMemcpyAsync(&ctx->h_res->err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return map_kernel_err(ctx->h_res->err);
}

// ---- block sort, then a coder, one call (include/b2rc.h)
uint64_t b2rc_blkrc_bound(int mode, uint64_t n, uint32_t block_size)
{
    const u64 inner = b2rc_bound(mode, b2rc_blk_encode_bound(n), block_size);
    return inner ? 16u + inner : 0u;
}

int b2rc_blkrc_encode_device(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                             uint8_t* d_dst, uint64_t dst_cap, uint64_t* out_n, void* cuda_stream)
{
    if(!ctx || !d_dst || (n && !d_src) || !mode_ok(mode) || !block_ok(block_size) || !aligned16(d_src) || !aligned16(d_dst)) {
        return B2RC_E_ARG;
    }
    if(out_n) {
        *out_n = b2rc_blkrc_bound(mode, n, block_size);
    }
    if(dst_cap < 16u + index_bytes(0)) {
        return B2RC_E_DST_SMALL;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const u64 nt = b2rc_blk_encode_bound(n);
    int rc;
    if((rc = grow(ctx, ctx->blk_tmp, ctx->blk_tmp_cap, (size_t)nt + 16)) != B2RC_OK) {
        return rc;
    }
    u64 made = 0;
    if((rc = b2rc_blk_encode_device(ctx, d_src, n, ctx->blk_tmp, nt, &made, st)) != B2RC_OK || made != nt) {
        return rc != B2RC_OK ? rc : B2RC_E_INTERNAL;
    }
    u64 inner = 0;
    rc = b2rc_encode_device(ctx, mode, block_size, ctx->blk_tmp, nt, d_dst + 16, dst_cap - 16, &inner, st);
    if(out_n) {
        *out_n = 16u + inner;
    }
    if(rc != B2RC_OK) {
        return rc;
    }
    u32* h = reinterpret_cast<u32*>(ctx->h_res->header);  // pinned
    h[0] = 0x53423242u;  // 'B','2','B','S'
    h[1] = 0;
    h[2] = (u32)n;
    h[3] = (u32)(n >> 32);
    CK(cudaMemcpyAsync(d_dst, h, 16, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return B2RC_OK;
}

int b2rc_blkrc_decode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                             uint64_t* out_n, void* cuda_stream)
{
    if(!ctx || !d_src || !aligned16(d_src) || !aligned16(d_dst)) {
        return B2RC_E_ARG;
    }
    if(n < 16u + B2RC_HEADER_BYTES + 8) {
        return B2RC_E_CORRUPT;
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CK(cudaMemcpyAsync(ctx->h_res->header, d_src, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    u32 h[4];
    memcpy(h, ctx->h_res->header, 16);
    const u64 orig = (u64)h[2] | ((u64)h[3] << 32);
    if(h[0] != 0x53423242u || h[1] != 0u || orig > (1ull << 46)) {
        return B2RC_E_CORRUPT;
    }
    if(out_n) {
        *out_n = orig;
    }
    if(dst_cap < orig || (orig && !d_dst)) {
        return B2RC_E_DST_SMALL;
    }
    const u64 nt = b2rc_blk_encode_bound(orig);
    int rc;
    if((rc = grow(ctx, ctx->blk_tmp, ctx->blk_tmp_cap, (size_t)nt + 16)) != B2RC_OK) {
        return rc;
    }
    u64 got = 0;
    if((rc = b2rc_decode_device(ctx, d_src + 16, n - 16, ctx->blk_tmp, nt, &got, st)) != B2RC_OK) {
        return rc == B2RC_E_DST_SMALL ? B2RC_E_CORRUPT : rc;  // the container disagrees with the size in front of it
    }
    if(got != nt) {
        return B2RC_E_CORRUPT;
    }
    u64 back = 0;
    if((rc = b2rc_blk_decode_device(ctx, ctx->blk_tmp, nt, d_dst, dst_cap, &back, st)) != B2RC_OK) {
        return rc;
    }
    return back == orig ? B2RC_OK : B2RC_E_CORRUPT;
}

// Host pointers: chunks of whole blocks, each on a stream of its own (copy in, kernel, copy out), so the
// copies of one chunk overlap the sort of another.  Pinned buffers give full overlap.
static int blk_host(b2rc_ctx* ctx, bool forward, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* out_n)
{
    if(!ctx || (n && (!src || !dst))) {
        return B2RC_E_ARG;
    }
    const u64 need = forward ? b2rc_blk_encode_bound(n) : b2rc_blk_decoded_size(n);
    if(out_n) {
        *out_n = need;
    }
    if(dst_cap < need) {
        return B2RC_E_DST_SMALL;
    }
    const u64 in_unit = forward ? BLK_N : BLK_CODED, out_unit = forward ? BLK_CODED : BLK_N;
    const u64 nb = n / in_unit;
    DeviceGuard g(ctx->device);
    int rc;
    if((rc = grow(ctx, ctx->stage_in, ctx->stage_in_cap, (size_t)(n + 16))) != B2RC_OK ||
       (rc = grow(ctx, ctx->stage_out, ctx->stage_out_cap, (size_t)(need + 16))) != B2RC_OK ||
       (forward && (rc = grow(ctx, ctx->blk_rounds, ctx->blk_rounds_cap, (size_t)(nb * 4 + 16))) != B2RC_OK) ||
       (forward && (rc = grow(ctx, ctx->blk_tie_list, ctx->blk_tie_cap, (size_t)(nb * 4 + 16))) != B2RC_OK)) {
        return rc;
    }
    if(forward) {
        ctx->blk_last_blocks = nb;
    }
    cudaStream_t s0 = ctx->pipe[0];
    if(forward) {
        CK(cudaMemsetAsync(ctx->blk_tie_list, 0, 4, s0));  // ordered before every chunk by index_ready
    }
    CK(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), s0));
    CK(cudaEventRecord(ctx->index_ready, s0));
    // chunk boundaries at even block numbers keep every chunk 4-byte aligned on the coded side
    u64 count = n / (16ull << 20);
    count = count < 1 ? 1 : (count > ctx->max_chunks ? ctx->max_chunks : count);
    u64 per = ((nb + count - 1) / count + 1) & ~1ull;
    per = per ? per : 2;
    u64 c = 0;
    for(u64 b0 = 0; b0 < nb; b0 += per, ++c) {
        const u64 b1 = b0 + per < nb ? b0 + per : nb;
        cudaStream_t st = ctx->pipe[1 + c % (B2RC_PIPE_STREAMS - 1)];
        CK(cudaStreamWaitEvent(st, ctx->index_ready, 0));
        CK(cudaMemcpyAsync(ctx->stage_in + b0 * in_unit, src + b0 * in_unit, (size_t)((b1 - b0) * in_unit), cudaMemcpyHostToDevice, st));
        rc = forward ? blk_forward_launch(ctx, ctx->stage_in, ctx->stage_out, b0, b1 - b0, st)
                     : blk_inverse_launch(ctx, ctx->stage_in, ctx->stage_out, b0, b1 - b0, st);
        if(rc != B2RC_OK) {
            return rc;
        }
        CK(cudaMemcpyAsync(dst + b0 * out_unit, ctx->stage_out + b0 * out_unit, (size_t)((b1 - b0) * out_unit), cudaMemcpyDeviceToHost, st));
    }
    if(n > nb * in_unit) {
        memcpy(dst + nb * out_unit, src + nb * in_unit, (size_t)(n - nb * in_unit));  // the tail never visits the device
    }
    for(int k = 1; k < B2RC_PIPE_STREAMS; ++k) {
        CK(cudaStreamSynchronize(ctx->pipe[k]));
    }
    if(forward) {
        std::vector<u64> fixed;
        if((rc = blk_fix_ties(ctx, ctx->stage_in, ctx->stage_out, nb, s0, &fixed)) != B2RC_OK) {
            return rc;
        }
        for(u64 b : fixed) {  // two bytes per replayed block
            CK(cudaMemcpyAsync(dst + b * BLK_CODED + BLK_N, ctx->stage_out + b * BLK_CODED + BLK_N, 2, cudaMemcpyDeviceToHost, s0));
        }
    }
    CK(cudaMemcpyAsync(&ctx->h_res->err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    return map_kernel_err(ctx->h_res->err);
}

int b2rc_blk_encode(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* out_n)
{
    return blk_host(ctx, true, src, n, dst, dst_cap, out_n);
}

int b2rc_blk_decode(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* out_n)
{
    return blk_host(ctx, false, src, n, dst, dst_cap, out_n);
}

int b2rc_blk_rounds(b2rc_ctx* ctx, uint32_t* rounds, uint64_t cap, uint64_t* nblocks)
{
    if(!ctx || (cap && !rounds)) {
        return B2RC_E_ARG;
    }
    if(nblocks) {
        *nblocks = ctx->blk_last_blocks;
    }
    const u64 k = cap < ctx->blk_last_blocks ? cap : ctx->blk_last_blocks;
    if(k) {
        DeviceGuard g(ctx->device);
        CK(cudaMemcpy(rounds, ctx->blk_rounds, (size_t)(k * 4), cudaMemcpyDeviceToHost));
    }
    return B2RC_OK;
}
}
