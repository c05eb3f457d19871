// The float screening of tlen_class (k_tilet, csrc/pss_kernels_tile.cu) against the exact forms it falls back to:
// whenever the screening accepts its estimate, the value must equal what the exact form returns -- mean
// rint(S / n) in doubles (opt.pyx:2166-2170), rms the integer decision of q_rootmean == rint(sqrt(Q / n)) in doubles
// (opt.pyx:2159-2163), std tlen_std from the exact integer sums -- and for std the exact form must not call the
// position ambiguous (there the tile is replayed in file order).  The kernel's approximate reciprocal / reciprocal
// square root are modelled as exact float operations perturbed so that the estimates are off by up to 0.8e-6 relative (the error bound is 5e-7, the margins allow 1e-6).
// Multisets: exact ties (values c, c, c, 0 x 6: std = c / 2), their neighbours, insert-size-like normals, mixtures with
// zeros and negatives, large n.   g++ -O2 scripts/tlen_fast_check.cpp && ./a.out
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <vector>

static int q_rootmean(unsigned long long sq, uint32_t n) { return (n == 0 || sq == 0) ? 0 : (int)rint(sqrt((double)sq / (double)n)); }
template <class U>
static int tlen_std_t(unsigned long long n, U ss, unsigned long long Q, bool *amb, float force) {
    const U nq = (U)n * (U)Q;
    const U A = nq >= ss ? nq - ss : 0;
    const U B = (U)n * (U)(n - 1);
    const double est = sqrt((double)A / (double)B);
    unsigned long long k = est > 0.5 ? (unsigned long long)(est - 0.5) : 0;
    while ((U)k * k * B > A) --k;
    while ((U)(k + 1) * (k + 1) * B <= A) ++k;
    const U T = (U)(2 * k + 1) * (2 * k + 1) * B, F = 4 * A;
    const U D = F > T ? F - T : T - F;
    const double thr = 256.0 * (double)n * (double)n * (double)Q * 1.1102230246251565e-16 + force;
    if (D == 0 || (double)D <= thr) *amb = true;
    return (int)(k + (F > T ? 1u : F == T ? (k & 1u) : 0u));
}
static int tlen_std(long long n, long long S, unsigned long long Q, bool *amb, float force) {
    if (n < 2) return 0;
    if (n < (1 << 20) && Q < (1ull << 40)) {
        const unsigned long long a = (unsigned long long)(S < 0 ? -S : S);
        return tlen_std_t<unsigned long long>((unsigned long long)n, a * a, Q, amb, force);
    }
    return tlen_std_t<unsigned __int128>((unsigned long long)n, (unsigned __int128)((__int128)S * (__int128)S), Q, amb, force);
}

static long n_fast[3], n_slow[3], bad;

static void check(unsigned n, long long S, unsigned long long Q, float noise) {
    if (n == 0) return;
    const bool small = n < 65536u && (Q >> 32) == 0 && (long long)(int)S == S;
    const float inv = (1.0f / (float)n) * (1.0f + noise);
    {
        const float e = (float)(int)S * inv * (1.0f + noise), r = rintf(e);
        if (small && fabsf(e - r) <= 0.4999f - 1e-6f * fabsf(e)) {
            ++n_fast[0];
            const int want = (int)rint((double)S / (double)n);
            if ((int)r != want) { if (bad++ < 10) printf("mean n %u S %lld: fast %d exact %d\n", n, S, (int)r, want); }
        } else ++n_slow[0];
    }
    if (Q != 0) {
        const float x = (float)(unsigned)Q * inv, e = sqrtf(x) * (1.0f + noise), r = rintf(e);
        if (small && fabsf(e - r) <= 0.4999f - 1e-6f * e) {
            ++n_fast[1];
            const int want = q_rootmean(Q, n);
            if ((int)r != want) { if (bad++ < 10) printf("rms n %u Q %llu: fast %d exact %d\n", n, Q, (int)r, want); }
        } else ++n_slow[1];
    }
    if (n < 2) return;
    {
        const unsigned long long nq = (unsigned long long)n * (unsigned)Q, ss = (unsigned long long)(S * S);
        const unsigned long long Aex = nq >= ss ? nq - ss : 0;
        const float v = ((float)Aex / (float)(n * (n - 1u))) * (1.0f + noise), e = sqrtf(v) * (1.0f + noise), r = rintf(e);
        if (small && (Aex == 0 || fabsf(e - r) <= 0.4995f - 1e-6f * e)) {
            ++n_fast[2];
            bool amb = false;
            const int want = tlen_std((long long)n, S, Q, &amb, 0.0f);
            const int got = Aex == 0 ? 0 : (int)r;
            if (got != want || amb) { if (bad++ < 10) printf("std n %u S %lld Q %llu: fast %d exact %d amb %d\n", n, S, Q, got, want, (int)amb); }
        } else ++n_slow[2];
    }
}

int main() {
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    auto gauss = [&]() { double a = 0; for (int i = 0; i < 12; ++i) a += (rnd() % 1000000) / 1e6; return a - 6.0; };
    auto run = [&](const std::vector<long long> &t) {
        long long S = 0; unsigned long long Q = 0;
        for (long long v : t) { S += v; Q += (unsigned long long)(v * v); }
        for (float noise : {0.0f, 0.3e-6f, -0.3e-6f}) check((unsigned)t.size(), S, Q, noise);
    };
    // exact ties of the standard deviation and their neighbours: m copies of c among n values, the rest 0 (and shifted)
    for (long long c = 1; c < 3000; ++c)
        for (int n : {9, 18, 2, 4, 50})
            for (long long d : {0ll, 1ll, -1ll, 7ll})
                for (long long shift : {0ll, 250ll, -400ll}) {
                    std::vector<long long> t(n, shift);
                    for (int i = 0; i < (n == 9 ? 3 : n / 2); ++i) t[i] = c + shift;
                    t[0] += d;
                    run(t);
                }
    // insert-size-like data, mixtures, large n
    for (int it = 0; it < 3000000; ++it) {
        const int kind = rnd() % 5;
        const int n = kind == 4 ? 1 + rnd() % 3000 : 1 + rnd() % 80;
        std::vector<long long> t(n);
        const double mu = (double)(rnd() % 900) - 200, sd = 1 + rnd() % 120;
        for (auto &v : t) {
            v = (long long)llround(mu + sd * gauss());
            if (kind == 1 && rnd() % 3 == 0) v = 0;
            if (kind == 2 && rnd() % 2) v = -v;
            if (kind == 3 && rnd() % 40 == 0) v = (long long)(rnd() % 200000) - 100000;      // a far mate
        }
        run(t);
    }
    for (int k = 0; k < 3; ++k) printf("%s: %ld accepted by the screening, %ld to the exact form (%.2f %%)\n", k == 0 ? "mean" : k == 1 ? "rms" : "std",
                                       n_fast[k], n_slow[k], 100.0 * n_slow[k] / (double)(n_fast[k] + n_slow[k]));
    printf("%ld bad\n", bad);
    return bad != 0;
}
