// Interval maxima over a tile of positions with O(1) updates: the max(mapq) of Mapq / MapqStrand
// (pysamstats/opt.pyx:919-925 `if mapq > self.max`) as a scatter instead of a per-position gather.
//
// A read contributes its value q to every position of an interval [a, b) of the tile.  The tile is cut
// into blocks of IMAX_BLK = 16 positions and the interval into at most three pieces:
//   head  [a, 16 ba)     a proper suffix of block ba - 1  -> one max into sfx[a]
//   body  blocks [ba, bb) whole blocks                     -> a sparse table over the blocks: the two
//                        (overlapping) power-of-two runs of blocks that cover it, tab[k][ba] and
//                        tab[k][bb - 2^k] with k = floor(log2(bb - ba))
//   tail  [16 bb, b)     a proper prefix of block bb      -> one max into pfx[b - 1]
// so a read costs at most FOUR atomic maxima whatever its length; an interval strictly inside one block
// (shorter than 16 positions and touching neither end of its block) updates its positions one by one in
// pt[].  After the last read the structure is resolved once per tile:
//   push-down  tab[k - 1][i] = max(tab[k - 1][i], tab[k][i], tab[k][i - 2^(k-1)])  for k = top .. 1,
//              which leaves the maximum over the body pieces covering block i in tab[0][i];
//   per block  result[p] = max(pt[p], tab[0][p >> 4], running max of sfx over the block's positions
//              <= p, running max of pfx over the block's positions >= p).
// Values are > 0 (a zero contributes nothing: the arrays start at zero, and so does the reference's max).
//
// The update and the resolution are written against plain arrays so that tests/cpp/imax_test.cpp can run
// exactly this code on the CPU against a brute-force maximum; the kernel passes atomicMax / warp shuffles.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define IMAX_HD __host__ __device__ __forceinline__
#else
#define IMAX_HD inline
#endif

#define IMAX_LOG 4
#define IMAX_BLK (1 << IMAX_LOG)

IMAX_HD int imax_floor_log2(uint32_t v) {       // v > 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// Where position p lives in sfx / pfx / pt: four words of padding after every 32 positions.  The resolution gives one
// lane a whole block (16 consecutive words, read as four 128-bit accesses); without the padding the lanes of a quarter
// warp would sit 16 words apart -- two bank groups for eight lanes, a four-way conflict on every access (ncu: 38 % of the
// kernel's shared-memory wavefronts) -- with it they spread over all eight.  A block stays contiguous and 16-byte aligned.
IMAX_HD int imax_idx(int p) { return p + ((p >> 5) << 2); }
constexpr int imax_padded(int n) { return n + n / 8; }      // words of an array for n positions

// number of table levels for a tile of `nblk` blocks (a power of two): levels 0 .. log2(nblk)
constexpr int imax_levels(int nblk) { return nblk <= 1 ? 1 : 1 + imax_levels(nblk >> 1); }

// One interval [a, b) (tile relative, 0 <= a < b <= tile length), value q > 0.
//   sfx, pfx, pt: the bank's arrays, position p at word imax_idx(p)
//   tab: the bank's level-0 row; level k lives at tab + k * tab_stride
//   amax(uint32_t *, uint32_t): atomic max
template <class AMax>
IMAX_HD void imax_update(int a, int b, uint32_t q, uint32_t *sfx, uint32_t *pfx, uint32_t *pt, uint32_t *tab,
                         int tab_stride, AMax amax) {
    const int ba = (a + IMAX_BLK - 1) >> IMAX_LOG, bb = b >> IMAX_LOG;      // whole blocks [ba, bb)
    if (ba <= bb) {
        // a <= 16 ba <= 16 bb <= b: head [a, 16 ba) and tail [16 bb, b) are a suffix / a prefix of their blocks
        if (a & (IMAX_BLK - 1)) amax(sfx + imax_idx(a), q);
        if (b & (IMAX_BLK - 1)) amax(pfx + imax_idx(b - 1), q);
        if (ba < bb) {
            const int nb = bb - ba, k = imax_floor_log2((uint32_t)nb);
            amax(tab + k * tab_stride + ba, q);
            if (nb != (1 << k)) amax(tab + k * tab_stride + bb - (1 << k), q);
        }
    } else {
        // strictly inside one block, touching neither of its ends
        for (int p = a; p < b; ++p) amax(pt + imax_idx(p), q);
    }
}

// Push-down of one bank's table, array form (the kernel does the same with one lane per block and
// shuffles): afterwards tab[0 * tab_stride + i] is the maximum over the body pieces covering block i.
inline void imax_pushdown_ref(uint32_t *tab, int tab_stride, int nblk) {
    for (int k = imax_levels(nblk) - 1; k >= 1; --k) {
        const int h = 1 << (k - 1);
        uint32_t *up = tab + k * tab_stride, *dn = tab + (k - 1) * tab_stride;
        for (int i = nblk - 1; i >= 0; --i) {
            uint32_t v = dn[i] > up[i] ? dn[i] : up[i];
            if (i >= h && up[i - h] > v) v = up[i - h];
            dn[i] = v;
        }
    }
}

// Resolution of one block (16 positions starting at sfx / pfx / pt) given the block's body maximum:
// the result replaces pt[].
IMAX_HD void imax_resolve_block(const uint32_t *sfx, const uint32_t *pfx, uint32_t *pt, uint32_t body) {
    uint32_t s[IMAX_BLK];
    uint32_t run = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < IMAX_BLK; ++i) { run = sfx[i] > run ? sfx[i] : run; s[i] = run; }
    run = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = IMAX_BLK - 1; i >= 0; --i) {
        run = pfx[i] > run ? pfx[i] : run;
        uint32_t v = pt[i] > body ? pt[i] : body;
        v = s[i] > v ? s[i] : v;
        pt[i] = run > v ? run : v;
    }
}
