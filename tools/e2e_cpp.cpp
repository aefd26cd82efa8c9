// e2e_cpp -- what a C++ caller of the drop-in classes gets, end to end, on the bench's stream:
// cpprcoder::RangeEncoder<>::encode / ::decode exactly as run_rangecoder calls them
// (test/main.cpp:270-290: a MemoryStream per direction, pageable memory everywhere), and the same
// through cpprcoder::PinnedStream (page-locked buffers).  Prints one JSON line; bench.py runs it.
//     e2e_cpp [bytes] [steps] [mode: 0 static, 1 adaptive]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../cpprcoder_b200/include/cpprcoder_b200.h"

using namespace cpprcoder;

// cpprcoder_b200/synth.py zipf(): splitmix64 in counter mode, four 16-bit draws per output, inverse CDF
static void zipf(u8* out, u64 n)
{
    double c[256], run = 0;
    for(int r = 0; r < 256; ++r) {
        run += 1.0 / (r + 1.0);
        c[r] = run;
    }
    std::vector<u8> lut(65536);
    int r = 0;
    for(int u = 0; u < 65536; ++u) {
        const double x = (u + 0.5) / 65536.0 * c[255];
        while(c[r] < x) {
            ++r;
        }
        lut[u] = (u8)r;
    }
    for(u64 i = 0; i < (n + 3) / 4; ++i) {
        u64 z = 0x5EEDC0DEull + (i + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        for(int k = 0; k < 4 && 4 * i + k < n; ++k) {
            out[4 * i + k] = lut[(z >> (16 * k)) & 0xFFFF];
        }
    }
}

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template<class Stream, class MakeEnc, class MakeDec>
static bool run(const char* name, int mode, const u8* src, u64 n, int steps, MakeEnc make_enc, MakeDec make_dec, double& gbps,
                double& enc_ms, double& dec_ms, u64& coded)
{
    double te = 0, td = 0;
    for(int it = -1; it < steps; ++it) {  // one untimed pass first
        Stream* enc = make_enc();
        Stream* dec = make_dec();
        const double t0 = now();
        bool ok;
        if(mode == 0) {
            RangeEncoder<Stream> e;
            ok = e.encode64(*enc, n, src);
            const double t1 = now();
            ok = ok && e.decode64(*dec, enc->size(), enc->get());
            const double t2 = now();
            if(it >= 0) {
                te += t1 - t0;
                td += t2 - t1;
            }
        } else {
            AdaptiveRangeEncoder<Stream> e;
            ok = e.initialize(*enc, (u32)n) && e.encode((s32)n, src).status_ == Status_Success;
            const double t1 = now();
            AdaptiveRangeDecoder<Stream> d;
            ok = ok && d.initialize(*dec) && d.decode((s32)enc->size(), enc->get()).status_ == Status_Success;
            const double t2 = now();
            if(it >= 0) {
                te += t1 - t0;
                td += t2 - t1;
            }
        }
        ok = ok && (u64)dec->size() == n && 0 == memcmp(dec->get(), src, n);
        coded = (u64)enc->size();
        delete enc;
        delete dec;
        if(!ok) {
            fprintf(stderr, "%s: round trip failed\n", name);
            return false;
        }
    }
    enc_ms = 1e3 * te / steps;
    dec_ms = 1e3 * td / steps;
    gbps = (double)n * steps / (te + td) / 1e9;
    return true;
}

int main(int argc, char** argv)
{
    const u64 n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 30);
    const int steps = argc > 2 ? atoi(argv[2]) : 3;
    const int mode = argc > 3 ? atoi(argv[3]) : 0;
    if(n == 0 || n > 0x7FFFFFF0ull || steps < 1) {
        fprintf(stderr, "usage: e2e_cpp [bytes < 2 GiB] [steps] [mode]\n");
        return 2;
    }
    std::vector<u8> src(n);
    zipf(src.data(), n);
    u64 sum = 0;
    for(u64 i = 0; i < n && i < (1u << 20); ++i) {
        sum += src[i];
    }
    double g0, e0, d0, g1, e1, d1;
    u64 coded = 0;
    const u64 cap = b2rc_bound(mode, n, B2RC_DEFAULT_BLOCK);
    if(!run<MemoryStream>("MemoryStream", mode, src.data(), n, steps, [&] { return new MemoryStream((s32)n); },
                          [&] { return new MemoryStream((s32)n); }, g0, e0, d0, coded)) {
        return 1;
    }
    // page-locked source as well: the caller who cares about the copies owns pinned memory throughout
    void* psrc = nullptr;
    if(B2RC_OK != b2rc_host_alloc(n, &psrc)) {
        return 1;
    }
    memcpy(psrc, src.data(), n);
    const bool ok = run<PinnedStream>("PinnedStream", mode, (const u8*)psrc, n, steps, [&] { return new PinnedStream(cap); },
                                      [&] { return new PinnedStream(n); }, g1, e1, d1, coded);
    b2rc_host_free(psrc);
    if(!ok) {
        return 1;
    }
    printf("{\"bytes\": %llu, \"steps\": %d, \"mode\": %d, \"first_mib_byte_sum\": %llu, \"container_bytes\": %llu, "
           "\"memory_stream\": {\"GBps\": %.4f, \"encode_ms\": %.3f, \"decode_ms\": %.3f, \"note\": \"run_rangecoder's calls: "
           "pageable source, a fresh MemoryStream per direction\"}, "
           "\"pinned_stream\": {\"GBps\": %.4f, \"encode_ms\": %.3f, \"decode_ms\": %.3f, \"note\": \"same calls, "
           "cpprcoder::PinnedStream and a page-locked source\"}}\n",
           (unsigned long long)n, steps, mode, (unsigned long long)sum, (unsigned long long)coded, g0, e0, d0, g1, e1, d1);
    return 0;
}
