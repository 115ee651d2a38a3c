// HOST helper of the upload path (see include/hfb200.h): multi-threaded copy of pageable host rows into a pinned staging
// buffer.  The reference consumes its NumPy arrays in place (PODProjector.py:726); a device path first has to get them across
// PCIe, and a cudaMemcpy from pageable memory is staged by the driver at ~10 GB/s.  torch's CPU copy_ reaches ~27 GB/s on the
// 16 cores of the B200 host; static partitioning over std::threads with non-temporal stores (no read-for-ownership of the
// destination lines) is ~1.8x faster, which moves the pageable upload from host-copy-bound towards PCIe-bound.
#include <emmintrin.h>

#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/hfb200.h"

namespace {

void copy_range(char* d, const char* s, size_t n) {
    if (reinterpret_cast<uintptr_t>(d) & 15) {
        memcpy(d, s, n);
        return;
    }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {  // one destination cache line per iteration, written around the cache
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
    if (i < n) memcpy(d + i, s + i, n - i);
    _mm_sfence();  // the DMA that follows must see the streamed lines
}

}  // namespace

extern "C" int hfb_host_copy(void* dst, const void* src, size_t bytes, int nthreads) {
    if (!dst || !src) return HFB_E_BADARG;
    if (bytes == 0) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (bytes < (size_t)(1 << 20)) nthreads = 1;  // not worth a thread below 1 MiB
    const size_t chunk = ((bytes / (size_t)nthreads) + 4095) & ~(size_t)4095;
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    std::vector<std::thread> workers;
    try {
        for (int t = 1; t < nthreads; ++t) {
            const size_t lo = (size_t)t * chunk;
            if (lo >= bytes) break;
            const size_t hi = lo + chunk < bytes ? lo + chunk : bytes;
            workers.emplace_back(copy_range, d + lo, s + lo, hi - lo);
        }
    } catch (...) {  // thread creation failed: the calling thread copies what is left
        for (auto& w : workers) w.join();
        const size_t done = (workers.size() + 1) * chunk;
        copy_range(d, s, chunk < bytes ? chunk : bytes);
        if (done < bytes) copy_range(d + done, s + done, bytes - done);
        return 0;
    }
    copy_range(d, s, chunk < bytes ? chunk : bytes);
    for (auto& w : workers) w.join();
    return 0;
}
