// Host-side unit test of the interval-maximum structure behind k_tilem (csrc/pss_imax.cuh): imax_update,
// the table push-down and imax_resolve_block -- the very functions the kernel calls -- against a brute-force
// per-position maximum, over random interval sets of every shape (whole tile, block aligned, inside one
// block, one position, clipped at either tile end), tile lengths 512 and 256.  Also runs the kept-gap
// enumeration of pss_lowq.cuh (min_baseq > 0: maximal runs of unmarked bases) against a bit-by-bit walk.
// Built and run by tests/test_host_logic.py.
#include "pss_imax.cuh"
#include "pss_lowq.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static uint32_t rnd() { static uint64_t s = 0x9E3779B97F4A7C15ull; s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }

static long test_tile(int TTN, int n_iv, int shape) {
    const int NB = TTN / IMAX_BLK, NLEV = imax_levels(NB);
    std::vector<uint32_t> sfx(imax_padded(TTN), 0), pfx(imax_padded(TTN), 0), pt(imax_padded(TTN), 0), tab((size_t)NLEV * NB, 0), want(TTN, 0);
    auto amax = [](uint32_t *p, uint32_t v) { if (v > *p) *p = v; };
    for (int i = 0; i < n_iv; ++i) {
        int a, b;
        switch (shape) {
        case 0: a = rnd() % TTN; b = a + 1 + rnd() % (TTN - a); break;                        // anything
        case 1: a = rnd() % TTN; b = a + 1 + rnd() % 15; if (b > TTN) b = TTN; break;          // short
        case 2: a = (rnd() % NB) * IMAX_BLK; b = a + IMAX_BLK * (1 + rnd() % (NB - a / IMAX_BLK)); break;   // whole blocks
        case 3: a = 0; b = 1 + rnd() % TTN; break;                                             // clipped on the left
        case 4: a = rnd() % TTN; b = TTN; break;                                               // clipped on the right
        case 5: a = rnd() % TTN; b = a + 100 + rnd() % 60; if (b > TTN) b = TTN; break;        // read shaped
        default: a = rnd() % TTN; b = a + 1; break;                                            // one position
        }
        const uint32_t q = 1 + rnd() % 255;
        imax_update(a, b, q, sfx.data(), pfx.data(), pt.data(), tab.data(), NB, amax);
        for (int p = a; p < b; ++p) if (q > want[p]) want[p] = q;
    }
    imax_pushdown_ref(tab.data(), NB, NB);
    for (int blk = 0; blk < NB; ++blk)
        imax_resolve_block(sfx.data() + imax_idx(blk * IMAX_BLK), pfx.data() + imax_idx(blk * IMAX_BLK), pt.data() + imax_idx(blk * IMAX_BLK), tab[blk]);
    long bad = 0;
    for (int p = 0; p < TTN; ++p) bad += pt[imax_idx(p)] != want[p];
    for (int p = 32; p < imax_padded(TTN); p += 36) for (int k = 0; k < 4; ++k) bad += (sfx[p + k] | pfx[p + k] | pt[p + k]) != 0;   // padding untouched
    return bad;
}

// kept gaps of one aligned span against a bit-by-bit walk of the marks
static long test_gaps(int iter) {
    long bad = 0;
    const int TTN = 512;
    for (int it = 0; it < iter; ++it) {
        const int lq = 1 + rnd() % 400;                      // read length
        const int nw = (lq + 31) / 32 + 1;
        std::vector<uint32_t> mw(nw, 0);
        const int mode = rnd() % 4;                          // no marks, sparse, dense, runs
        for (int i = 0; i < lq; ++i) {
            bool m = mode == 0 ? false : mode == 1 ? rnd() % 20 == 0 : mode == 2 ? rnd() % 2 == 0 : ((i / (1 + it % 37)) & 1);
            if (m) mw[i >> 5] |= 1u << (i & 31);
        }
        const int y = rnd() % lq, l = 1 + rnd() % (lq - y);  // the span: query [y, y + l)
        const int t0 = 10000, x = t0 - 300 + (int)(rnd() % 900);    // reference start, anywhere around the tile
        std::vector<int> want(TTN, 0), got(TTN, 0);
        for (int j = 0; j < l; ++j) {
            const int p = x + j - t0, qi = y + j;
            if (p >= 0 && p < TTN && !((mw[qi >> 5] >> (qi & 31)) & 1)) want[p] = 1;
        }
        int last_end = -1;
        lowq_gaps(mw.data(), x, l, y, t0, TTN, [&](int p, int len) {
            if (len <= 0 || p < 0 || p + len > TTN || p <= last_end) ++bad;      // maximal runs, in order, never adjacent
            last_end = p + len;
            for (int k = 0; k < len; ++k) got[p + k] += 1;
        });
        for (int p = 0; p < TTN; ++p) bad += got[p] != want[p];
    }
    return bad;
}

int main() {
    long total = 0;
    for (int TTN : {512, 256})
        for (int shape = 0; shape < 7; ++shape)
            for (int n_iv : {1, 2, 5, 40, 300}) {
                long bad = 0;
                for (int rep = 0; rep < 200; ++rep) bad += test_tile(TTN, n_iv, shape);
                if (bad) printf("tile %d shape %d intervals %d: %ld wrong positions\n", TTN, shape, n_iv, bad);
                total += bad;
            }
    const long gbad = test_gaps(20000);
    if (gbad) printf("gaps: %ld wrong\n", gbad);
    total += gbad;
    printf("imax_test: %s\n", total ? "FAILED" : "ok");
    return total ? 1 : 0;
}
