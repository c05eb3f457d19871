#include <cstdio>
#include <cstdint>
#include <cmath>
#include <initializer_list>
static int ref(unsigned long long sq, uint32_t n) { return n ? (int)rint(sqrt((double)sq / (double)n)) : 0; }
static int mine(unsigned long long sq, uint32_t n, float eps) {
    if (n == 0 || sq == 0) return 0;
    float x = (float)sq / (float)n;
    if ((sq >> 44) || x > 1.0e9f) return ref(sq, n);
    uint32_t r = (uint32_t)(sqrtf(x) * (1.0f + eps) + 0.5f);
    const unsigned long long f = 4 * sq;
    const uint32_t a = 2 * r + 1;
    const unsigned long long hi = (unsigned long long)(a * a) * n;
    if (f > hi || (f == hi && (r & 1u))) return (int)(r + 1);
    if (r > 0) {
        const uint32_t b = 2 * r - 1;
        const unsigned long long lo = (unsigned long long)(b * b) * n;
        if (f < lo || (f == lo && (r & 1u))) return (int)(r - 1);
    }
    return (int)r;
}
int main() {
    long bad = 0, cnt = 0;
    for (float eps : {0.0f, 3e-6f, -3e-6f})
    for (uint32_t n = 1; n <= 1500; ++n)
        for (uint32_t k = 0; k <= 300; ++k) {
            unsigned long long c = (unsigned long long)(2 * k + 1) * (2 * k + 1) * n;
            for (long long sq = (long long)(c / 4) - 2; sq <= (long long)(c / 4) + 2; ++sq) {
                if (sq < 0) continue;
                ++cnt;
                if (ref(sq, n) != mine(sq, n, eps)) { if (bad < 10) printf("bad sq=%lld n=%u ref=%d mine=%d\n", sq, n, ref(sq, n), mine(sq, n, eps)); ++bad; }
            }
        }
    unsigned long long x = 88172645463325252ull;
    for (long i = 0; i < 100000000; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        uint32_t n = (uint32_t)(x % 70000) + 1;
        unsigned long long sq = (x >> 20) % ((unsigned long long)n * 65025ull + 1);
        ++cnt;
        float eps = (i & 1) ? 2e-6f : -2e-6f;
        if (ref(sq, n) != mine(sq, n, eps)) { if (bad < 10) printf("bad sq=%llu n=%u\n", sq, n); ++bad; }
    }
    // large tlen-like values
    for (long i = 0; i < 50000000; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        uint32_t n = (uint32_t)(x % 1000) + 1;
        unsigned long long sq = (x >> 22) % (1ull << 43);
        ++cnt;
        if (ref(sq, n) != mine(sq, n, (i & 1) ? 2e-6f : -2e-6f)) { if (bad < 10) printf("bad L sq=%llu n=%u ref=%d mine=%d\n", sq, n, ref(sq,n), mine(sq,n,0)); ++bad; }
    }
    printf("checked %ld, bad %ld\n", cnt, bad);
    return bad != 0;
}
