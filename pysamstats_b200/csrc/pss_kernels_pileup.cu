// The pileup kernel: per-position statistics for one contig, all 15 PileupStat classes of
// the reference (pysamstats/opt.pyx:107-1519) fused with their rec() finalisation.
//
// Design (B200-first): the reference walks htslib's list of live reads once
// per column.  Here every thread OWNS one genome position and every warp owns 32
// consecutive positions; the tile's candidate reads (records, packed bases, qualities, tlen)
// are staged in shared memory by bulk async copies, and the warp iterates, in file order,
// over the reads that overlap its 32 positions (a contiguous index range, because reads are
// start-sorted) and walks each read's CIGAR warp-uniformly.  All counters of a position live in that thread's
// registers, so there are no atomics and no inter-thread traffic at all, the column order
// of the reference (needed for the order-dependent Welford std of tlen) is preserved, and
// a warp's stores are 128-byte coalesced rows of the SoA output.  Hot counters are packed
// 8-bit lanes (one 32-bit add updates four counters) spilled to 32-bit registers every 255
// reads, so any depth is exact.
#include "pss_internal.cuh"

#define FULL 0xffffffffu

// ------------------------------------------------------------------ numerics (A.5)
// numpy.around == round-half-even == rint(); IEEE double sqrt / div are correctly rounded.
__device__ __forceinline__ int rootmean(unsigned long long sq, long long n) {
    return n > 0 ? (int)rint(sqrt((double)sq / (double)n)) : 0;   // opt.pyx:2159-2163
}
__device__ __forceinline__ int meani(long long s, long long n) {
    return n > 0 ? (int)rint((double)s / (double)n) : 0;          // opt.pyx:2166-2170
}

// 4-bit code whose letter in "=ACMGRSVTWYHKDBN" (opt.pyx:72) equals the upper-cased
// reference character; 16 = no such code (never equal to a read base).
__device__ __forceinline__ int ref_code(int c) {
    switch (c) {
    case '=': return 0; case 'A': return 1; case 'C': return 2; case 'M': return 3;
    case 'G': return 4; case 'R': return 5; case 'S': return 6; case 'V': return 7;
    case 'T': return 8; case 'W': return 9; case 'Y': return 10; case 'H': return 11;
    case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15;
    default: return 16;
    }
}

// spread the low 4 bits of m to the low bit of 4 bytes
__device__ __forceinline__ uint32_t spread4(uint32_t m) { return ((m & 15u) * 0x00204081u) & 0x01010101u; }

// unpack 4 x 8-bit lanes into 32-bit accumulators
__device__ __forceinline__ void unpack4(uint32_t v, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
    a += v & 255u; b += (v >> 8) & 255u; c += (v >> 16) & 255u; d += v >> 24;
}

struct ReadCtx {       // warp-uniform per-read values
    uint32_t cls;
    int tlen;
};

// Every accumulator offers:
//   NCELL          1, 2 (strand) or 4 (strand x proper-pair) register banks
//   QUAL / REF     needs base qualities / the reference base
//   begin()        per-read, warp-uniform
//   match<C>()     a non-deletion event of this thread's position in bank C
//   del<C>()       a deletion / ref-skip event
//   flush()        spill packed lanes (called at most every 255 reads)
//   store()        finalise (rec()) and write this position's row
// Events arrive already filtered (min_baseq / no_del / read-level filters).

// ---- coverage, coverage_strand, coverage_gc (opt.pyx:107-211, :1471-1519) -------------
template <int STAT>
struct AccCoverage {
    static const int NCELL = 4;
    static const bool QUAL = false, REF = false;
    uint32_t n[4];
    __device__ void init() { n[0] = n[1] = n[2] = n[3] = 0; }
    __device__ void setup(const uint2 *) {}
    __device__ void begin(const ReadCtx &) {}
    template <int C> __device__ void match(bool act, int, int, bool, int) { n[C] += act; }
    template <int C> __device__ void del(bool act, bool) { n[C] += act; }
    __device__ void flush() {}
    __device__ void store(int32_t *o, int64_t st) {
        // bank = rev | pp << 1
        uint32_t fwd = n[0] + n[2], rev = n[1] + n[3], pp = n[2] + n[3];
        if (STAT == PSS_COVERAGE) { o[0] = fwd + rev; o[st] = pp; }
        else if (STAT == PSS_COVERAGE_GC) { o[st] = fwd + rev; o[2 * st] = pp; }   // field 0 = gc (k_gc)
        else { o[0] = fwd + rev; o[st] = fwd; o[2 * st] = rev; o[3 * st] = pp; o[4 * st] = n[2]; o[5 * st] = n[3]; }
    }
};

// ---- coverage_ext(_strand) (opt.pyx:220-426): 8 read classes x strand ------------------
template <int STAT>
struct AccCoverageExt {
    static const int NCELL = 2;   // bank = rev
    static const bool QUAL = false, REF = false;
    uint32_t lo[2], hi[2];        // packed: lo = all,pp,mun,other ; hi = same,face,soft,dup
    uint32_t w[2][8];
    uint32_t ilo, ihi;            // this read's increments (uniform)
    __device__ void init() {
        lo[0] = lo[1] = hi[0] = hi[1] = 0;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) w[s][k] = 0;
    }
    __device__ void setup(const uint2 *) {}
    __device__ void begin(const ReadCtx &r) {
        uint32_t m = (r.cls >> CLS_EXT_SHIFT) & 255u;
        ilo = spread4(m); ihi = spread4(m >> 4);
    }
    template <int C> __device__ void match(bool act, int, int, bool, int) { if (act) { lo[C] += ilo; hi[C] += ihi; } }
    template <int C> __device__ void del(bool act, bool) { if (act) { lo[C] += ilo; hi[C] += ihi; } }
    __device__ void flush() {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            unpack4(lo[s], w[s][0], w[s][1], w[s][2], w[s][3]);
            unpack4(hi[s], w[s][4], w[s][5], w[s][6], w[s][7]);
            lo[s] = hi[s] = 0;
        }
    }
    __device__ void store(int32_t *o, int64_t st) {
        if (STAT == PSS_COVERAGE_EXT) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k * st] = w[0][k] + w[1][k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                o[(3 * k) * st] = w[0][k] + w[1][k]; o[(3 * k + 1) * st] = w[0][k]; o[(3 * k + 2) * st] = w[1][k];
            }
        }
    }
};

// ---- variation(_strand) (opt.pyx:435-670) ---------------------------------------------
// per bank: lo = A,C,G,T   hi = N, matches, non-deletion events, insertions   ev = all, deletions
__constant__ uint2 c_var_lut[17] = {
    {0u, 0x00010000u}, {0x00000001u, 0x00010000u}, {0x00000100u, 0x00010000u}, {0u, 0x00010000u},
    {0x00010000u, 0x00010000u}, {0u, 0x00010000u}, {0u, 0x00010000u}, {0u, 0x00010000u},
    {0x01000000u, 0x00010000u}, {0u, 0x00010000u}, {0u, 0x00010000u}, {0u, 0x00010000u},
    {0u, 0x00010000u}, {0u, 0x00010000u}, {0u, 0x00010000u}, {0u, 0x00010001u}, {0u, 0x00010000u}};

template <int STAT>
struct AccVariation {
    static const int NCELL = 4;
    static const bool QUAL = false, REF = true;
    uint32_t lo[4], hi[4], ev[4];
    uint32_t w[4][10];   // all, match, nondel, del, ins, A, C, T, G, N
    const uint2 *lut;
    __device__ void init() {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            lo[c] = hi[c] = ev[c] = 0;
#pragma unroll
            for (int k = 0; k < 10; ++k) w[c][k] = 0;
        }
    }
    __device__ void setup(const uint2 *l) { lut = l; }
    __device__ void begin(const ReadCtx &) {}
    template <int C> __device__ void match(bool act, int code, int, bool ins, int refcode) {
        if (act) {
            uint2 inc = lut[code];
            lo[C] += inc.x;
            hi[C] += inc.y + ((code == refcode && code < 16) ? 0x100u : 0u) + (ins ? 0x01000000u : 0u);
            ev[C] += 1u;
        }
    }
    template <int C> __device__ void del(bool act, bool refskip) {
        if (act) ev[C] += refskip ? 1u : 0x101u;     // opt.pyx:486-488
    }
    __device__ void flush() {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t a = 0, cc = 0, g = 0, t = 0;
            unpack4(lo[c], a, cc, g, t);
            w[c][5] += a; w[c][6] += cc; w[c][8] += g; w[c][7] += t;
            uint32_t nn = 0, m = 0, nd = 0, in = 0;
            unpack4(hi[c], nn, m, nd, in);
            w[c][9] += nn; w[c][1] += m; w[c][2] += nd; w[c][4] += in;
            w[c][0] += ev[c] & 255u; w[c][3] += (ev[c] >> 8) & 255u;
            lo[c] = hi[c] = ev[c] = 0;
        }
    }
    __device__ void store(int32_t *o, int64_t st) {
        // output groups: reads, matches, mismatches, deletions, insertions, A, C, T, G, N
#pragma unroll
        for (int g = 0; g < 10; ++g) {
            uint32_t v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t x;
                switch (g) {
                case 0: x = w[c][0]; break;
                case 1: x = w[c][1]; break;
                case 2: x = w[c][2] - w[c][1]; break;      // mismatches = non-deletions - matches
                case 3: x = w[c][3]; break;
                case 4: x = w[c][4]; break;
                default: x = w[c][g]; break;
                }
                v[c] = x;
            }
            uint32_t fwd = v[0] + v[2], rev = v[1] + v[3], pp = v[2] + v[3];
            if (STAT == PSS_VARIATION) { o[(2 * g) * st] = fwd + rev; o[(2 * g + 1) * st] = pp; }
            else {
                int32_t *q = o + (int64_t)(6 * g) * st;
                q[0] = fwd + rev; q[st] = fwd; q[2 * st] = rev; q[3 * st] = pp; q[4 * st] = v[2]; q[5 * st] = v[3];
            }
        }
    }
};

// ---- mapq(_strand) (opt.pyx:912-1084) ---------------------------------------------------
template <int STAT>
struct AccMapq {
    static const int NCELL = 4;
    static const bool QUAL = false, REF = false;
    uint32_t n[4], nz[4], mx[4];
    unsigned long long sq[4];
    uint32_t q;
    __device__ void init() {
#pragma unroll
        for (int c = 0; c < 4; ++c) { n[c] = nz[c] = mx[c] = 0; sq[c] = 0; }
    }
    __device__ void setup(const uint2 *) {}
    __device__ void begin(const ReadCtx &r) { q = (r.cls >> CLS_MAPQ_SHIFT) & 255u; }
    template <int C> __device__ void ev(bool act) {
        if (act) { n[C]++; nz[C] += (q == 0); mx[C] = max(mx[C], q); sq[C] += (unsigned long long)(q * q); }
    }
    template <int C> __device__ void match(bool act, int, int, bool, int) { ev<C>(act); }
    template <int C> __device__ void del(bool act, bool) { ev<C>(act); }
    __device__ void flush() {}
    __device__ void store(int32_t *o, int64_t st) {
        // helper order all, fwd, rev, pp, pp_fwd, pp_rev from banks rev | pp << 1
        uint32_t N[6] = {n[0] + n[1] + n[2] + n[3], n[0] + n[2], n[1] + n[3], n[2] + n[3], n[2], n[3]};
        uint32_t Z[6] = {nz[0] + nz[1] + nz[2] + nz[3], nz[0] + nz[2], nz[1] + nz[3], nz[2] + nz[3], nz[2], nz[3]};
        unsigned long long S[6] = {sq[0] + sq[1] + sq[2] + sq[3], sq[0] + sq[2], sq[1] + sq[3], sq[2] + sq[3], sq[2], sq[3]};
        uint32_t M[6] = {max(max(mx[0], mx[1]), max(mx[2], mx[3])), max(mx[0], mx[2]), max(mx[1], mx[3]),
                         max(mx[2], mx[3]), mx[2], mx[3]};
        if (STAT == PSS_MAPQ) {
            o[0] = N[0]; o[st] = N[3]; o[2 * st] = Z[0]; o[3 * st] = Z[3];
            o[4 * st] = rootmean(S[0], N[0]); o[5 * st] = rootmean(S[3], N[3]);
            o[6 * st] = M[0]; o[7 * st] = M[3];
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                o[k * st] = N[k]; o[(6 + k) * st] = Z[k];
                o[(12 + k) * st] = rootmean(S[k], N[k]); o[(18 + k) * st] = M[k];
            }
        }
    }
};

// ---- baseq(_strand) and baseq_ext(_strand) (opt.pyx:1093-1462) ---------------------------
template <int STAT>
struct AccBaseq {
    static const int NCELL = 4;
    static const bool EXT = (STAT == PSS_BASEQ_EXT || STAT == PSS_BASEQ_EXT_STRAND);
    static const bool STRAND = (STAT == PSS_BASEQ_STRAND || STAT == PSS_BASEQ_EXT_STRAND);
    static const bool QUAL = true, REF = EXT;
    uint32_t n[4], nnd[4], nm[4];
    unsigned long long sq[4], sqm[4];
    __device__ void init() {
#pragma unroll
        for (int c = 0; c < 4; ++c) { n[c] = nnd[c] = nm[c] = 0; sq[c] = sqm[c] = 0; }
    }
    __device__ void setup(const uint2 *) {}
    __device__ void begin(const ReadCtx &) {}
    template <int C> __device__ void match(bool act, int code, int q, bool, int refcode) {
        if (act) {
            unsigned long long s = (unsigned long long)(q * q);
            n[C]++; nnd[C]++; sq[C] += s;
            if (EXT && code == refcode && code < 16) { nm[C]++; sqm[C] += s; }
        }
    }
    template <int C> __device__ void del(bool act, bool) { n[C] += act; }   // counted, no quality (opt.pyx:1108-1112)
    __device__ void flush() {}
    template <class T> __device__ static void six(const T *b, T *o6) {
        o6[0] = b[0] + b[1] + b[2] + b[3]; o6[1] = b[0] + b[2]; o6[2] = b[1] + b[3];
        o6[3] = b[2] + b[3]; o6[4] = b[2]; o6[5] = b[3];
    }
    __device__ void store(int32_t *o, int64_t st) {
        uint32_t N[6], D[6], Mn[6];
        unsigned long long S[6], Sm[6];
        six(n, N); six(nnd, D); six(sq, S);
        if (EXT) { six(nm, Mn); six(sqm, Sm); }
        if (STAT == PSS_BASEQ) {
            o[0] = N[0]; o[st] = N[3]; o[2 * st] = rootmean(S[0], D[0]); o[3 * st] = rootmean(S[3], D[3]);
        } else if (STAT == PSS_BASEQ_STRAND) {
#pragma unroll
            for (int k = 0; k < 6; ++k) { o[k * st] = N[k]; o[(6 + k) * st] = rootmean(S[k], D[k]); }
        } else if (STAT == PSS_BASEQ_EXT) {
            // reads, matches, mismatches (x all,pp) then rms for the same three groups
            const int idx[2] = {0, 3};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int k = idx[j];
                o[j * st] = N[k]; o[(2 + j) * st] = Mn[k]; o[(4 + j) * st] = D[k] - Mn[k];
                o[(6 + j) * st] = rootmean(S[k], D[k]); o[(8 + j) * st] = rootmean(Sm[k], Mn[k]);
                o[(10 + j) * st] = rootmean(S[k] - Sm[k], D[k] - Mn[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                o[k * st] = N[k]; o[(6 + k) * st] = Mn[k]; o[(12 + k) * st] = D[k] - Mn[k];
                o[(18 + k) * st] = rootmean(S[k], D[k]); o[(24 + k) * st] = rootmean(Sm[k], Mn[k]);
                o[(30 + k) * st] = rootmean(S[k] - Sm[k], D[k] - Mn[k]);
            }
        }
    }
};

// ---- tlen(_strand) (opt.pyx:679-903) --------------------------------------------------
// TlenHelper keeps Welford's running mean / M2 in doubles in column (= file) order; this
// thread sees its column in exactly that order, so the recurrence is replayed verbatim
// (explicit _rn intrinsics: no FMA contraction, like the reference's compiled C).
struct Welford {
    long long n, s, sq;
    double m, m2;
    __device__ void init() { n = s = sq = 0; m = m2 = 0.0; }
    __device__ void update(bool act, long long x) {
        if (act) {
            n++; s += x; sq += x * x;
            double d = __dsub_rn((double)x, m);
            m = __dadd_rn(m, __ddiv_rn(d, (double)n));
            double d2 = __dsub_rn((double)x, m);
            m2 = __dadd_rn(m2, __dmul_rn(d, d2));
        }
    }
    __device__ int std() const { return n < 2 ? 0 : (int)rint(sqrt(__ddiv_rn(m2, (double)(n - 1)))); }
};

template <int STAT>
struct AccTlen {
    static const int NCELL = 1;
    static const bool STRAND = (STAT == PSS_TLEN_STRAND);
    static const bool QUAL = false, REF = false;
    uint32_t nfwd, nrev;
    Welford h[STRAND ? 6 : 2];    // strand: all,fwd,rev,pp,pp_fwd,pp_rev ; else all(paired), pp
    bool rev, pp, paired;
    long long t;
    __device__ void init() {
        nfwd = nrev = 0;
#pragma unroll
        for (int k = 0; k < (STRAND ? 6 : 2); ++k) h[k].init();
    }
    __device__ void setup(const uint2 *) {}
    __device__ void begin(const ReadCtx &r) {
        rev = r.cls & CLS_REV; pp = r.cls & CLS_PP; paired = r.cls & CLS_PAIRED; t = r.tlen;
    }
    __device__ void ev(bool act) {
        if (rev) nrev += act; else nfwd += act;
        if (STRAND) {
            if (paired) { h[0].update(act, t); if (rev) h[2].update(act, t); else h[1].update(act, t); }
            if (pp) { h[3].update(act, t); if (rev) h[5].update(act, t); else h[4].update(act, t); }
        } else {
            if (paired) h[0].update(act, t);
            if (pp) h[1].update(act, t);
        }
    }
    template <int C> __device__ void match(bool act, int, int, bool, int) { ev(act); }
    template <int C> __device__ void del(bool act, bool) { ev(act); }
    __device__ void flush() {}
    __device__ void store(int32_t *o, int64_t st) {
        if (!STRAND) {
            o[0] = nfwd + nrev; o[st] = (int)h[0].n; o[2 * st] = (int)h[1].n;
            o[3 * st] = meani(h[0].s, h[0].n); o[4 * st] = meani(h[1].s, h[1].n);
            o[5 * st] = rootmean((unsigned long long)h[0].sq, h[0].n);
            o[6 * st] = rootmean((unsigned long long)h[1].sq, h[1].n);
            o[7 * st] = h[0].std(); o[8 * st] = h[1].std();
        } else {
            o[0] = nfwd + nrev; o[st] = nfwd; o[2 * st] = nrev;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                o[(3 + k) * st] = (int)h[k].n;
                o[(9 + k) * st] = meani(h[k].s, h[k].n);
                o[(15 + k) * st] = rootmean((unsigned long long)h[k].sq, h[k].n);
                o[(21 + k) * st] = h[k].std();
            }
        }
    }
};

// ------------------------------------------------------------------- the read walk
//
// seq / qual point at the read's first base (shared-memory staging buffers in the normal
// kernel, global memory in the long-read variant); all lanes of a warp walk the same read.

template <class ACC, int C>
__device__ __forceinline__ void match_span(ACC &acc, int min_baseq, int p, int x, int l, int y, bool ins_end,
                                           const uint8_t *seq, const uint8_t *qual, bool hasseq, int refcode) {
    int off = p - x;
    bool act = (unsigned)off < (unsigned)l;
    int code = 16, q = 0;
    if (act) {
        int qpos = y + off;
        if (hasseq) code = (seq[qpos >> 1] >> ((~qpos & 1) << 2)) & 15;                 // opt.pyx:2154
        if (ACC::QUAL || qual) {
            if (qual && hasseq) q = qual[qpos];   // no SEQ / QUAL: quality 0 (the reference reads past the record here)
            if (min_baseq > 0 && q < min_baseq) act = false;                            // opt.pyx:1900-1903
        }
    }
    acc.template match<C>(act, code, q, ins_end && off == l - 1, refcode);
}

template <class ACC, int C>
__device__ __forceinline__ void process_read(ACC &acc, const PileupArgs &a, const RInfo &r, int tlen, const uint8_t *seq,
                                             const uint8_t *qual, int p, int w0, int w1, int refcode) {
    ReadCtx rc;
    rc.cls = r.cls;
    rc.tlen = tlen;
    acc.begin(rc);
    const bool hasseq = r.cls & CLS_HASSEQ;
    if (r.cls & CLS_SIMPLE) {
        match_span<ACC, C>(acc, a.min_baseq, p, r.pos, r.end - r.pos, 0, false, seq, qual, hasseq, refcode);
        return;
    }
    if (r.cls & CLS_NOREF) {   // mapped, CIGAR consumes no reference: one deletion-like event at pos
        if (!a.no_del) acc.template del<C>(p == r.pos, false);
        return;
    }
    const uint32_t *cg = a.cigar + r.cigar_off;
    const int nc = (int)r.n_cigar;
    int x = r.pos, y = 0;
    for (int k = 0; k < nc && x < w1; ++k) {
        uint32_t w = __ldg(cg + k);
        int op = w & 15, l = (int)(w >> 4);
        if (op_is_match(op)) {
            if (x + l > w0) {
                bool ins = false;   // htslib `indel > 0` at the last base of the op (A.1)
                if (k + 1 < nc && x + l - 1 >= w0 && x + l - 1 < w1) {
                    uint32_t w2 = __ldg(cg + k + 1);
                    int op2 = w2 & 15;
                    if (op2 == OP_I) ins = (w2 >> 4) > 0;
                    else if (op2 == OP_P && k + 2 < nc) {
                        long long l3 = 0;
                        for (int j = k + 2; j < nc; ++j) {
                            uint32_t w3 = __ldg(cg + j);
                            int op3 = w3 & 15;
                            if (op3 == OP_I) l3 += (w3 >> 4);
                            else if (op3 == OP_D || op3 == OP_M || op3 == OP_N || op3 == OP_EQ || op3 == OP_X) break;
                        }
                        ins = l3 > 0;
                    }
                }
                match_span<ACC, C>(acc, a.min_baseq, p, x, l, y, ins, seq, qual, hasseq, refcode);
            }
            x += l; y += l;
        } else if (op == OP_D || op == OP_N) {
            if (x + l > w0 && !a.no_del)                                              // opt.pyx:1904-1906
                acc.template del<C>((unsigned)(p - x) < (unsigned)l, op == OP_N);
            x += l;
        } else if (op == OP_I || op == OP_S) y += l;
    }
}

template <class ACC>
__device__ __forceinline__ void dispatch_read(ACC &acc, const PileupArgs &a, const RInfo &r, int tlen, const uint8_t *seq,
                                              const uint8_t *qual, int p, int w0, int w1, int refcode) {
    if (ACC::NCELL == 4) {
        switch (r.cls & 3u) {
        case 0: process_read<ACC, 0>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode); break;
        case 1: process_read<ACC, 1>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode); break;
        case 2: process_read<ACC, 2>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode); break;
        default: process_read<ACC, 3>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode); break;
        }
    } else if (ACC::NCELL == 2) {
        if (r.cls & CLS_REV) process_read<ACC, 1>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode);
        else process_read<ACC, 0>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode);
    } else process_read<ACC, 0>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode);
}

// staging buffers of the normal kernel (reads up to ~150 bp fill a whole chunk; longer reads
// shrink the chunk; a read that does not fit at all needs PSS_FLAG_LONG_READS)
#define GCH 96
#define GSEQ_BYTES (GCH * 76 + 64)
#define GQUAL_BYTES (GCH * 152 + 64)

struct GatherSmem {
    RInfo ri[PSS_CHUNK];                         // GCH used when staged
    long long qoff[GCH + 4];
    int32_t tlen[GCH + 8];
    __align__(16) uint8_t seq[GSEQ_BYTES + 32];
    __align__(16) uint8_t qual[GQUAL_BYTES + 32];
};

// STAGED: RInfo, sequence, qualities, quality offsets and tlen of the tile's candidate reads
// are staged in shared memory by bulk async copies (cp.async.bulk + mbarrier); each warp then
// binary-searches its first relevant read.  !STAGED: sequence / qualities are read from
// global memory (any read length).
// the accumulators the replay pass behind k_tilet uses (the only ones instantiated with LIST)
template <class A> struct IsTlenAcc { static const bool V = false; };
template <int STAT> struct IsTlenAcc<AccTlen<STAT>> { static const bool V = true; };

// LIST: the replay pass (a.redo: a list of tiles walked by a small fixed grid).  A compile-time switch: an ordinary launch
// has one tile per CTA and its loop below is visibly a single round, so the kernels of the other statistics keep the code
// (and the register allocation) they had before the replay list existed.
template <class ACC, bool STAGED, bool LIST = false>
__global__ void __launch_bounds__(PSS_TILE)
k_pileup(PileupArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    GatherSmem &S = *reinterpret_cast<GatherSmem *>(smem_raw);
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_n;
    __shared__ long long s_seqlo, s_quallo, s_qofflo, s_tlenlo;
    // per-base increment table in shared memory: the index diverges across lanes, which the
    // constant cache would serialise
    __shared__ uint2 s_lut[17];

    const int tid = threadIdx.x;
    // The replay pass behind k_tilet (a.redo != NULL) runs on a small fixed grid: a.redo[0] = number of flagged
    // 256-position tiles, a.redo[1..] = their indices, CTA b takes entries b, b + gridDim.x, ...  (One CTA per tile of the
    // range, returning at once for the unflagged ones, cost 0.75 ms per 50 Mb in block launches alone -- a quarter of
    // tlen_strand.)  An ordinary launch has one tile per CTA and runs the loop body once.
    const int n_mine = LIST ? a.redo[0] : (int)gridDim.x;
    if (tid < 17) s_lut[tid] = c_var_lut[tid];
    uint32_t parity = 0;          // every bulk-copy phase issued below is waited for exactly once: the parity carries over
    bool first = true;
    for (int it = blockIdx.x; it < n_mine; it += gridDim.x) {
    const int tile = LIST ? a.redo[1 + it] : it;
    const int64_t t0 = a.start + (int64_t)tile * PSS_TILE;
    const int64_t pl = t0 + tid;
    const bool valid = pl < a.end;
    const int p = (int)pl;
    const int w0 = (int)(t0 + (tid & ~31)), w1 = w0 + 32;
    // (replay: the first eight bytes of the 48-byte descriptor of the 512-position tile this one is half of)
    const int2 rng = LIST ? a.tile_rng[(size_t)(tile >> 1) * 6] : a.tile_rng[tile];
    const bool need_seq = ACC::REF, need_qual = a.qual != nullptr, need_tlen = a.tlen != nullptr;

    // one chunk: thread 0 sizes it and issues the bulk copies
    auto issue = [&](int cb) {
        int n = min(GCH, rng.y - cb);
        uint32_t b1 = 0, b2 = 0, b3 = 0, b4 = 0;
        const uint8_t *p1 = nullptr, *p2 = nullptr, *p3 = nullptr, *p4 = nullptr;
        long long slo = 0, shi = 0, qlo = 0, qhi = 0;
        while (n > 0) {
            if (need_seq) { slo = a.seq_off[cb]; shi = a.seq_off[cb + n]; }
            if (need_qual) { qlo = a.qual_off[cb]; qhi = a.qual_off[cb + n]; }
            if (shi - slo + 32 <= GSEQ_BYTES && qhi - qlo + 32 <= GQUAL_BYTES) break;
            if (n == 1) { atomicOr(a.toolong, 1); n = 0; }
            else n >>= 1;
        }
        if (n > 0) {
            if (need_seq) stage_span(a.seq4, slo, shi, 1, &p1, &b1, &s_seqlo);
            if (need_qual) {
                stage_span(a.qual, qlo, qhi, 1, &p2, &b2, &s_quallo);
                stage_span(a.qual_off, (long long)cb, (long long)cb + n, 8, &p3, &b3, &s_qofflo);
            }
            if (need_tlen) stage_span(a.tlen, (long long)cb, (long long)cb + n, 4, &p4, &b4, &s_tlenlo);
        }
        s_n = n;
        fence_proxy_async();
        if (n > 0) {
            mbar_arrive_expect_tx(&s_bar, (uint32_t)n * (uint32_t)sizeof(RInfo) + b1 + b2 + b3 + b4);
            bulk_g2s(S.ri, a.ri + cb, (uint32_t)n * (uint32_t)sizeof(RInfo), &s_bar);
            if (b1) bulk_g2s(S.seq, p1, b1, &s_bar);
            if (b2) bulk_g2s(S.qual, p2, b2, &s_bar);
            if (b3) bulk_g2s(S.qoff, p3, b3, &s_bar);
            if (b4) bulk_g2s(S.tlen, p4, b4, &s_bar);
        } else mbar_arrive_expect_tx(&s_bar, 0);
    };

    if (!first) __syncthreads();  // later rounds of the replay: everyone is done with the staging buffers
    if (STAGED && tid == 0) {     // thread 0 starts the bulk copies of the first chunk before anything else happens
        if (first) mbar_init(&s_bar, 1);
        if (rng.x < rng.y) issue(rng.x);
    }
    first = false;
    ACC acc;
    acc.init();
    acc.setup(s_lut);
    int refcode = 16;
    if (ACC::REF && a.ref && valid) {
        int64_t ro = pl - a.ref_start;
        if (ro >= 0 && ro < a.ref_len) {
            int c = a.ref[ro];
            if (c >= 'a' && c <= 'z') c -= 32;
            refcode = ref_code(c);
        }
    }
    __syncthreads();
    const int maxspan = a.st_ro->maxspan;
    bool raw = false;
    bool done = false;
    int nproc = 0;
    for (int base = rng.x; base < rng.y;) {
        int cnt;
        if (STAGED) {
            mbar_wait(&s_bar, parity);
            parity ^= 1;
            cnt = s_n;
            if (cnt == 0) break;
        } else {
            cnt = min(PSS_CHUNK, rng.y - base);
            __syncthreads();
            for (int t = tid; t < cnt * 2; t += PSS_TILE)
                reinterpret_cast<uint4 *>(S.ri)[t] = __ldg(reinterpret_cast<const uint4 *>(a.ri + base) + t);
            __syncthreads();
        }
        if (!done) {
            // first read of the chunk that can reach this warp's window: pos >= w0 - maxspan + 1
            int j = 0;
            {
                int lo = 0, hi = cnt;
                const int key = w0 - maxspan + 1;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (S.ri[mid].pos < key) lo = mid + 1; else hi = mid; }
                j = lo;
            }
            const long long seqlo = STAGED ? s_seqlo : 0, quallo = STAGED ? s_quallo : 0;
            const int qskew = STAGED ? (int)(base - s_qofflo) : 0, tskew = STAGED ? (int)(base - s_tlenlo) : 0;
            for (; j < cnt; ++j) {
                const uint4 h = *reinterpret_cast<const uint4 *>(&S.ri[j]);     // pos, end, cls, n_cigar
                const int rpos = (int)h.x, rend = (int)h.y;
                const uint32_t cls = h.z;
                if (rpos >= w1) { done = true; break; }
                if (rend <= w0 || !(cls & CLS_ADMITTED)) continue;
                raw |= (unsigned)(p - rpos) < (unsigned)(rend - rpos);
                if (!(cls & CLS_COUNTED)) continue;
                const RInfo r = S.ri[j];
                int tlen = 0;
                const uint8_t *seq, *qual = nullptr;
                if (STAGED) {
                    seq = S.seq + (r.seq_off - seqlo);
                    if (need_qual) qual = S.qual + (S.qoff[j + qskew] - quallo);
                    if (need_tlen) tlen = S.tlen[j + tskew];
                } else {
                    const int64_t ridx = (int64_t)base + j;
                    seq = a.seq4 + r.seq_off;
                    if (need_qual) qual = a.qual + __ldg(a.qual_off + ridx);
                    if (need_tlen) tlen = __ldg(a.tlen + ridx);
                }
                dispatch_read<ACC>(acc, a, r, tlen, seq, qual, p, w0, w1, refcode);
                if (++nproc == 255) { acc.flush(); nproc = 0; }
            }
        }
        base += cnt;
        if (STAGED && base < rng.y) {
            __syncthreads();                       // staging buffers free again
            if (tid == 0) issue(base);
        }
    }
    acc.flush();
    if (valid) {
        acc.store(a.out + (pl - a.start), a.out_stride);
        a.rowmask[pl - a.start] = raw ? 1 : 0;
    }
    if (!LIST) break;
    }
}

template <class ACC>
static cudaError_t launch_one(const PileupArgs &a, int64_t ntiles, bool staged, cudaStream_t s) {
    const size_t smem = sizeof(GatherSmem);
    static bool configured_dev[64] = {false};     // per instantiation and device (attributes are per device)
    int dev__ = 0;
    cudaGetDevice(&dev__);
    bool &configured = configured_dev[dev__ & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_pileup<ACC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_pileup<ACC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (a.redo) {
        if constexpr (IsTlenAcc<ACC>::V) {
            static bool configured_list[64] = {false};
            if (!configured_list[dev__ & 63]) {
                cudaError_t e = cudaFuncSetAttribute(k_pileup<ACC, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                configured_list[dev__ & 63] = true;
            }
            k_pileup<ACC, true, true><<<(unsigned)ntiles, PSS_TILE, smem, s>>>(a);
            return cudaGetLastError();
        } else return cudaErrorInvalidValue;
    }
    if (staged) k_pileup<ACC, true><<<(unsigned)ntiles, PSS_TILE, smem, s>>>(a);
    else k_pileup<ACC, false><<<(unsigned)ntiles, PSS_TILE, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pileup(int stat, const PileupArgs &a, int64_t ntiles, bool staged, cudaStream_t s, int64_t *launches) {
    if (ntiles <= 0) return cudaSuccess;
    if (launches) ++*launches;
    switch (stat) {
    case PSS_COVERAGE: return launch_one<AccCoverage<PSS_COVERAGE>>(a, ntiles, staged, s);
    case PSS_COVERAGE_STRAND: return launch_one<AccCoverage<PSS_COVERAGE_STRAND>>(a, ntiles, staged, s);
    case PSS_COVERAGE_GC: return launch_one<AccCoverage<PSS_COVERAGE_GC>>(a, ntiles, staged, s);
    case PSS_COVERAGE_EXT: return launch_one<AccCoverageExt<PSS_COVERAGE_EXT>>(a, ntiles, staged, s);
    case PSS_COVERAGE_EXT_STRAND: return launch_one<AccCoverageExt<PSS_COVERAGE_EXT_STRAND>>(a, ntiles, staged, s);
    case PSS_VARIATION: return launch_one<AccVariation<PSS_VARIATION>>(a, ntiles, staged, s);
    case PSS_VARIATION_STRAND: return launch_one<AccVariation<PSS_VARIATION_STRAND>>(a, ntiles, staged, s);
    case PSS_TLEN: return launch_one<AccTlen<PSS_TLEN>>(a, ntiles, staged, s);
    case PSS_TLEN_STRAND: return launch_one<AccTlen<PSS_TLEN_STRAND>>(a, ntiles, staged, s);
    case PSS_MAPQ: return launch_one<AccMapq<PSS_MAPQ>>(a, ntiles, staged, s);
    case PSS_MAPQ_STRAND: return launch_one<AccMapq<PSS_MAPQ_STRAND>>(a, ntiles, staged, s);
    case PSS_BASEQ: return launch_one<AccBaseq<PSS_BASEQ>>(a, ntiles, staged, s);
    case PSS_BASEQ_STRAND: return launch_one<AccBaseq<PSS_BASEQ_STRAND>>(a, ntiles, staged, s);
    case PSS_BASEQ_EXT: return launch_one<AccBaseq<PSS_BASEQ_EXT>>(a, ntiles, staged, s);
    case PSS_BASEQ_EXT_STRAND: return launch_one<AccBaseq<PSS_BASEQ_EXT_STRAND>>(a, ntiles, staged, s);
    default: return cudaErrorInvalidValue;
    }
}
