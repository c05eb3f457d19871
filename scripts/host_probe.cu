// Host-side limits of the e2e path on the GPU box: pinned D2H / H2D bandwidth (alone and together)
// and the widening throughput of N host threads (u16 -> i32, u8 -> i32, non-temporal stores).
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
__attribute__((target("avx2"))) static void widen16(const uint16_t *src, int32_t *dst, size_t n) {
    for (size_t i = 0; i + 16 <= n; i += 16) {
        __m256i a = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(src + i)));
        __m256i b = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(src + i + 8)));
        _mm256_stream_si256((__m256i *)(dst + i), a);
        _mm256_stream_si256((__m256i *)(dst + i + 8), b);
    }
    _mm_sfence();
}
__attribute__((target("avx2"))) static void widen8(const uint8_t *src, int32_t *dst, size_t n) {
    for (size_t i = 0; i + 16 <= n; i += 16) {
        __m128i v = _mm_loadu_si128((const __m128i *)(src + i));
        __m256i a = _mm256_cvtepu8_epi32(v);
        __m256i b = _mm256_cvtepu8_epi32(_mm_srli_si128(v, 8));
        _mm256_stream_si256((__m256i *)(dst + i), a);
        _mm256_stream_si256((__m256i *)(dst + i + 8), b);
    }
    _mm_sfence();
}
int main(int argc, char **argv) {
    int nth = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
    const size_t N = (size_t)1 << 30;   // elements
    char *d0, *d1, *h0, *h1;
    cudaMalloc(&d0, N * 2); cudaMalloc(&d1, N * 2);
    cudaMallocHost(&h0, N * 2); cudaMallocHost(&h1, N * 2);
    cudaMemset(d0, 1, N * 2); memset(h1, 1, N * 2);
    cudaStream_t sa, sb; cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now(); cudaMemcpyAsync(h0, d0, N * 2, cudaMemcpyDeviceToHost, sa); cudaStreamSynchronize(sa); double t1 = now();
        cudaMemcpyAsync(d1, h1, N * 2, cudaMemcpyHostToDevice, sb); cudaStreamSynchronize(sb); double t2 = now();
        cudaMemcpyAsync(h0, d0, N * 2, cudaMemcpyDeviceToHost, sa); cudaMemcpyAsync(d1, h1, N * 2, cudaMemcpyHostToDevice, sb);
        cudaStreamSynchronize(sa); cudaStreamSynchronize(sb); double t3 = now();
        if (rep) printf("D2H %.1f GB/s, H2D %.1f GB/s, both at once %.1f + %.1f GB/s\n", N * 2 / (t1 - t0) / 1e9, N * 2 / (t2 - t1) / 1e9,
                        N * 2 / (t3 - t2) / 1e9, N * 2 / (t3 - t2) / 1e9);
    }
    int32_t *wide = (int32_t *)aligned_alloc(4096, N * 4);
    memset(wide, 0, N * 4);
    for (int mode = 0; mode < 2; ++mode)
        for (int threads : {nth / 4 > 0 ? nth / 4 : 1, nth / 2 > 0 ? nth / 2 : 1, nth}) {
            for (int rep = 0; rep < 2; ++rep) {
                double t0 = now();
                std::vector<std::thread> th;
                for (int t = 0; t < threads; ++t)
                    th.emplace_back([=] {
                        size_t per = (N / threads) & ~(size_t)63, o = per * t;
                        if (mode == 0) widen16((const uint16_t *)h0 + o, wide + o, per); else widen8((const uint8_t *)h0 + o, wide + o, per);
                    });
                for (auto &x : th) x.join();
                double dt = now() - t0;
                if (rep) printf("widen u%d -> i32, %d threads: %.1f G elements/s (%.1f GB/s written)\n", mode ? 8 : 16, threads, N / dt / 1e9, N * 4 / dt / 1e9);
            }
        }
    // widening while a D2H copy is in flight (what the pipeline does)
    {
        double t0 = now();
        cudaMemcpyAsync(h1, d0, N * 2, cudaMemcpyDeviceToHost, sa);
        std::vector<std::thread> th;
        for (int t = 0; t < nth; ++t)
            th.emplace_back([=] { size_t per = (N / nth) & ~(size_t)63, o = per * t; widen8((const uint8_t *)h0 + o, wide + o, per); });
        for (auto &x : th) x.join();
        double t1 = now();
        cudaStreamSynchronize(sa);
        double t2 = now();
        printf("concurrent: widen u8 %.1f G elements/s while D2H ran at %.1f GB/s\n", N / (t1 - t0) / 1e9, N * 2 / (t2 - t0) / 1e9);
    }
    return 0;
}
