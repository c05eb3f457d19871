// The logic of mq_root32 / mq_root_tie (k_tilem, csrc/pss_kernels_tile.cu) against the double formula of the reference
// (rootmean, opt.pyx:2159-2163): every boundary neighbourhood (sq around (r +- 1/2)^2 n) for n up to 65535 and 3e7 random
// cases, each with the float estimate perturbed both ways.  g++ -O2 scripts/mq_root_check.cpp && ./a.out
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
static uint32_t mq_root_tie(uint32_t sq, uint32_t n, uint32_t r, bool above) {
    const uint32_t c = above ? 2 * r + 1 : 2 * r - 1;
    const unsigned long long cn = (unsigned long long)(c * c) * n;
    const uint32_t t = (uint32_t)(cn >> 2);
    const bool exact = ((uint32_t)cn & 3u) == 0, odd = r & 1u, eq = sq == t;
    const bool up = above & ((sq > t) | (eq & exact & odd));
    const bool dn = !above & ((sq < t) | (eq & (!exact | odd)));
    return r + (up ? 1u : 0u) - (dn ? 1u : 0u);
}
// mq_root32 of csrc/pss_kernels_tile.cu; the kernel's approximate division / reciprocal square root are modelled as the
// exact float operations perturbed by `noise` relative (they are good to 2 ulp each: < 1e-6 together)
static uint32_t mq_root32(uint32_t sq, uint32_t n, float noise) {
    const float x = (float)sq / (float)n;
    const float rf = sqrtf(x) * (1.0f + noise);
    const float rr = rintf(rf);
    uint32_t r = (uint32_t)rr;
    if (fabsf(rf - rr) > 0.499f) r = mq_root_tie(sq, n, r, rf >= rr);
    return sq ? r : 0u;
}
int main() {
    long bad = 0, cases = 0;
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    auto check = [&](uint32_t sq, uint32_t n) {
        const uint32_t want = (uint32_t)rint(sqrt((double)sq / (double)n));     // numpy.around = half to even = rint
        for (float noise : {0.0f, 3e-6f, -3e-6f, 1e-6f, -1e-6f}) { ++cases; if (mq_root32(sq, n, noise) != want) { if (bad < 10) printf("sq %u n %u want %u got %u\n", sq, n, want, mq_root32(sq, n, noise)); ++bad; } }
    };
    // every boundary neighbourhood: sq around (r +- 0.5)^2 n for every n up to 65535 (step) and every r
    for (uint32_t n = 1; n <= 65535; n += (n < 600 ? 1 : 37))
        for (uint32_t r = 0; r <= 255; ++r)
            for (int side = -1; side <= 1; side += 2) {
                const double b = (r + 0.5 * side) * (r + 0.5 * side) * n;
                if (b < 0 || b > 255.0 * 255.0 * n + 2) continue;
                for (long d = -2; d <= 2; ++d) { const double v = floor(b) + d; if (v >= 0 && v <= 255.0 * 255.0 * n) check((uint32_t)v, n); }
            }
    for (int i = 0; i < 30000000; ++i) { const uint32_t n = 1 + rnd() % 65535; const uint32_t sq = (uint32_t)(rnd() % ((uint64_t)65025 * n + 1)); check(sq, n); }
    printf("%ld cases, %ld bad\n", cases, bad);
    return bad != 0;
}
