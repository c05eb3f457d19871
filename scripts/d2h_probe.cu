// D2H bandwidth probe: 1-D vs pitched 2-D copies with the geometry pss_pileup_host_multi uses.
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t nf = 60, cn = 1118208, npos = 24999936;      // one sub-chunk of variation_strand
    size_t dev_bytes = nf * cn * 4, host_bytes = nf * npos * 4;
    char *d, *h;
    cudaMalloc(&d, dev_bytes);
    cudaMallocHost(&h, host_bytes);
    cudaMemset(d, 1, dev_bytes);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            double t0 = now();
            size_t total = 0;
            for (size_t cs = 0; cs + cn <= npos; cs += cn) {
                if (mode == 0) cudaMemcpyAsync(h + cs * 4 * nf, d, dev_bytes, cudaMemcpyDeviceToHost, s);
                else if (mode == 1) cudaMemcpy2DAsync(h + cs * 4, npos * 4, d, cn * 4, cn * 4, nf, cudaMemcpyDeviceToHost, s);
                else for (size_t f = 0; f < nf; ++f) cudaMemcpyAsync(h + (f * npos + cs) * 4, d + f * cn * 4, cn * 4, cudaMemcpyDeviceToHost, s);
                total += dev_bytes;
            }
            cudaStreamSynchronize(s);
            double dt = now() - t0;
            if (rep) printf("mode %d (%s): %.1f GB/s\n", mode, mode == 0 ? "1-D contiguous" : mode == 1 ? "2-D pitched" : "one 1-D copy per row", total / dt / 1e9);
        }
    }
    return 0;
}
