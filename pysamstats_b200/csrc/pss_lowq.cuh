// min_baseq > 0: walking the mark bits k_lowq wrote (one bit per read base, set when the base's quality is below
// the threshold -- the test of stat_pileup, pysamstats/opt.pyx:1900-1903).  Host + device, so that
// tests/cpp/imax_test.cpp runs the same code against a bit-by-bit walk.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LOWQ_HD __host__ __device__ __forceinline__
#else
#define LOWQ_HD inline
#endif

LOWQ_HD uint32_t lowq_load(const uint32_t *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
LOWQ_HD int lowq_ffs(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __ffs((int)v);
#else
    return __builtin_ffs((int)v);
#endif
}

// Calls f(p, len) for every MAXIMAL run of unmarked bases of the aligned span [x, x + l) (first base query
// index y; mw = the read's mark words) that lies in the tile [t0, t0 + tt): p is the tile-relative position of
// the run's first base.  Runs come in increasing order and are never adjacent (a run that crosses a 32-bit mark
// word is merged), so a read without marked bases is ONE call however long it is.
template <class F>
LOWQ_HD void lowq_gaps(const uint32_t *mw, int x, int l, int y, int t0, int tt, F f) {
    const int a = x - t0 > 0 ? x - t0 : 0, b = x + l - t0 < tt ? x + l - t0 : tt;
    if (a >= b) return;
    const int qa = y + (a + t0 - x), qb = y + (b + t0 - x);       // query range [qa, qb)
    const int w0 = qa >> 5, w1 = (qb - 1) >> 5;
    int rs = -1, re = -1;                                         // the open run [rs, re)
    for (int w = w0; w <= w1; ++w) {
        uint32_t m = ~lowq_load(mw + w);
        const int base = w << 5;
        if (base < qa) m &= 0xffffffffu << (qa - base);
        if (base + 32 > qb) m &= 0xffffffffu >> (base + 32 - qb);
        while (m) {
            const int bit = lowq_ffs(m) - 1;
            const uint32_t inv = ~(m >> bit);
            const int len = inv ? lowq_ffs(inv) - 1 : 32 - bit;
            const int p = a + (base + bit - qa);
            if (p == re) re += len;
            else {
                if (rs >= 0) f(rs, re - rs);
                rs = p; re = p + len;
            }
            m = (len + bit >= 32) ? 0u : (m & (0xffffffffu << (bit + len)));
        }
    }
    if (rs >= 0) f(rs, re - rs);
}
