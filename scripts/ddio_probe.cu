// Does a small, cache-resident staging ring make the D2H + widen pipeline cheaper than a large one?
// T threads, each: D2H of one slice of its pinned ring (own stream) -> sync -> widen u8 -> i32 into a
// large output.  Ring of R slices per thread; total staging = T * R * slice bytes.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <immintrin.h>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__attribute__((target("avx2"))) static void widen8(const uint8_t *src, int32_t *dst, size_t n) {
    for (size_t i = 0; i + 16 <= n; i += 16) {
        __m128i v = _mm_loadu_si128((const __m128i *)(src + i));
        _mm256_stream_si256((__m256i *)(dst + i), _mm256_cvtepu8_epi32(v));
        _mm256_stream_si256((__m256i *)(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(v, 8)));
    }
    _mm_sfence();
}
int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 16;
    const size_t total = (size_t)4 << 30;          // elements widened per configuration
    uint8_t *dev;
    cudaMalloc(&dev, 1 << 28);
    cudaMemset(dev, 3, 1 << 28);
    int32_t *wide = (int32_t *)aligned_alloc(4096, ((size_t)1 << 30) * 4);
    memset(wide, 0, ((size_t)1 << 30) * 4);
    for (size_t slice : {(size_t)256 << 10, (size_t)1 << 20, (size_t)4 << 20, (size_t)32 << 20})
        for (int ring : {2, 8}) {
            if (slice * ring * T > ((size_t)3 << 30)) continue;
            uint8_t *stage;
            cudaMallocHost(&stage, slice * ring * T);
            memset(stage, 1, slice * ring * T);
            const size_t per_thread = total / T, iters = per_thread / slice;
            double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([=] {
                    cudaStream_t s;
                    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
                    cudaEvent_t ev[8];
                    for (int r = 0; r < ring; ++r) cudaEventCreateWithFlags(&ev[r], cudaEventDisableTiming);
                    uint8_t *mine = stage + (size_t)t * ring * slice;
                    int32_t *out = wide + ((size_t)t << 26);          // 64 Mi elements of output per thread, reused
                    // keep ring - 1 copies in flight ahead of the widening
                    for (size_t k = 0; k < iters + ring - 1; ++k) {
                        if (k < iters) {
                            int r = (int)(k % ring);
                            cudaMemcpyAsync(mine + (size_t)r * slice, dev + ((k * slice) & ((1 << 27) - 1)), slice, cudaMemcpyDeviceToHost, s);
                            cudaEventRecord(ev[r], s);
                        }
                        if (k + 1 >= (size_t)ring) {
                            size_t j = k + 1 - ring;
                            int r = (int)(j % ring);
                            cudaEventSynchronize(ev[r]);
                            widen8(mine + (size_t)r * slice, out + ((j * slice) & (((size_t)1 << 26) - 1)), slice);
                        }
                    }
                    cudaStreamSynchronize(s);
                });
            for (auto &x : th) x.join();
            double dt = now() - t0;
            printf("slice %6zu KiB x ring %d x %d threads (staging %5zu MiB): %.1f G elements/s\n", slice >> 10, ring, T,
                   (slice * ring * T) >> 20, (double)(iters * slice * T) / dt / 1e9);
            cudaFreeHost(stage);
        }
    return 0;
}
