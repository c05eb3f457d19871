// The tile kernel: the per-read scatter path for the counting statistics (coverage,
// coverage_strand, coverage_gc, coverage_ext, coverage_ext_strand, variation,
// variation_strand -- pysamstats/opt.pyx:107-670, :1471-1519) when min_baseq == 0.
//
// Design (sm_100a):
//   * one CTA owns a tile of TT genome positions; its per-position counters live in shared
//     memory for the whole life of the tile.  A counter word holds TWO 16-bit counters of
//     the same position (forward | reverse strand << 16), so variation_strand needs 20 word
//     arrays (40 KB for 512 positions) and three CTAs fit on an SM;
//   * the tile's candidate reads are a contiguous index range of the start-sorted batch;
//     their 32-byte records, packed 4-bit sequences, CIGAR words and the tile of the 4-bit
//     packed reference are staged into shared memory by bulk async copies (cp.async.bulk ->
//     UBLKCP, completion on an mbarrier) issued before anything else happens, chunk by
//     chunk for arbitrarily deep tiles;
//   * reads are walked once.  Everything that is constant along an aligned span (read
//     counts per strand / proper-pair / mate class, deletions) is an interval add: +1 at
//     the span start, -1 at its end in a difference array.  Bases are NOT visited one by
//     one: up to four threads share a read, each XORs eight 4-bit codes (one 32-bit word)
//     against the reference tile packed the same way, and only the mismatching bases (<1 %)
//     become point updates.  matches and the A/C/G/T/N counters follow exactly from
//     matches = aligned - mismatches and the reference code of the position (App. C);
//   * after the last chunk the difference arrays are prefix-summed in place (warp shuffle
//     scan over the packed words: the halves cannot borrow from each other in the final
//     sums) and the tile is finalised: every field of these statistics is linear in the
//     counters; each thread produces four consecutive positions of a subset of the fields
//     and writes them as 128-bit stores (512 contiguous bytes per warp and field);
//   * 16-bit halves hold at most 65535: a tile is processed in passes of at most 65535
//     reads, later passes ADD their (linear) results to the output, so any depth is exact.
// Shared-memory atomics are few (a handful per read), so the kernel is bound by the output
// stream to HBM.
#include "pss_internal.cuh"
#include "pss_imax.cuh"
#include "pss_lowq.cuh"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define FULL 0xffffffffu

#define TT PSS_TTILE          // positions per tile
// variation_strand: 256 threads x FOUR CTAs per SM.  The 20 counter arrays are 40 KB; with a 136-read staging chunk, 192
// staged CIGAR words and the row mask's own counter array only when a read-level filter needs it, a CTA takes 56.8 KB
// of shared memory and a fourth fits (227 KB per SM).  Measured on 100 Mb (kernel ms): 320 threads x 3 CTAs, 192-read
// chunks 5.00 (0.80 of the HBM roofline); 256 x 4 with 128-read chunks 4.83, with 136-read chunks 4.80 (0.84); 224 x 4
// 5.23; 192 x 4 5.08; 320 x 3 with 128-read chunks 5.71 -- the store phase of one CTA hides behind the walks of more others.
#ifndef TTHREADS
#define TTHREADS 256
#endif
#ifndef TCTAS
#define TCTAS 4               // CTAs per SM the variation kernels are compiled for
#endif
// Coverage kinds: little work per tile, so more, smaller CTAs per SM.  Measured per 50 Mb (threads x CTAs the kernel is
// compiled for; coverage / coverage_strand / coverage_gc | coverage_ext / coverage_ext_strand, ms; profiles/
// r02b_cov_launch_shapes.txt): 192 x 8 (rounds 1-2a) 0.217 / 0.300 / 0.306 | 0.449 / 0.819; 256 x 6 0.260 / 0.316 / 0.349 |
// 0.486 / 0.861; 160 x 10 0.183 / 0.303 / 0.273 | 0.491 / 1.016; 128 x 10 0.198 / 0.265 / 0.287 | 0.396 / 0.786; 128 x 12
// 0.190 / 0.255 / 0.278 | 0.458 / 0.891; 128 x 14 0.175 / 0.358 / 0.270 (36 registers: coverage_strand's six columns
// spill) ; 96 x 16 0.176 / 0.273 / 0.265 | 0.418 / 0.816.  So: 128 threads; 12 CTAs for the COV kind, 14 for its
// two-column statistics (coverage, coverage_gc), 10 for the EXT kind.
#ifndef TTHREADS_COV
#define TTHREADS_COV 128
#endif
#ifndef TCTAS_COV
#define TCTAS_COV 12
#endif
#ifndef TCTAS_COV2
#define TCTAS_COV2 14         // coverage, coverage_gc
#endif
#ifndef TTHREADS_EXT
#define TTHREADS_EXT 128
#endif
#ifndef TCTAS_EXT
#define TCTAS_EXT 10
#endif
template <int KIND> struct KindLaunch { static const int NT = TTHREADS_COV, CTAS = TCTAS_COV; };
template <> struct KindLaunch<1 /* KIND_EXT */> { static const int NT = TTHREADS_EXT, CTAS = TCTAS_EXT; };
template <> struct KindLaunch<2 /* KIND_VAR */> { static const int NT = TTHREADS, CTAS = TCTAS; };
// variation (no strand split): half the counter arrays (45 KB of shared memory per CTA), so five CTAs fit an SM; measured
// per 50 Mb (threads / CTAs the kernel is compiled for): 384/3 1.96, 320/3 1.85, 320/4 1.72, 256/4 1.68, 256/5 1.58,
// 192/6 1.51 ms -- more, smaller CTAs overlap the phases of their tiles better
#ifndef TTHREADS_VAR1
#define TTHREADS_VAR1 192
#endif
#ifndef TCTAS_VAR1
#define TCTAS_VAR1 6
#endif
template <> struct KindLaunch<6 /* KIND_VAR1 */> { static const int NT = TTHREADS_VAR1, CTAS = TCTAS_VAR1; };
// CTAs per SM a (kind, statistic) is compiled for
template <int KIND, int STAT> struct StatCtas { static const int V = KindLaunch<KIND>::CTAS; };
template <> struct StatCtas<0 /* KIND_COV */, PSS_COVERAGE> { static const int V = TCTAS_COV2; };
template <> struct StatCtas<0 /* KIND_COV */, PSS_COVERAGE_GC> { static const int V = TCTAS_COV2; };
#define NQUAD (TT / 4)        // finalisation: quads of positions
#ifndef TCH
#define TCH 136               // reads staged per chunk, variation kinds (a 512-position tile of 30x / 150 bp has ~134 candidates);
#endif                        // also the first chunk the tile descriptors describe
#ifndef TCH_COV
#define TCH_COV 128           // coverage kinds (records only: 4 KB; one staged read per thread)
#endif
#define TSEQ_BYTES (TCH * 76 + 64)   // staged sequence bytes per chunk (150 bp reads fit a full chunk)
#define REFN_WORDS (TT / 8)
#ifndef TCIG_WORDS
#define TCIG_WORDS 192        // staged CIGAR words per chunk (else the chunk's reads load them from global)
#endif
#define PASS_READS 65535      // reads per pass: what a 16-bit counter can hold
#ifndef TFIN_LOOP
#define TFIN_LOOP 1           // variation_strand: the finalisation as a rolled loop over the field groups
#endif
#ifndef TSTCS
#define TSTCS 1               // int32 result columns are written once and never read again: streaming stores (13.20 -> 13.04 ms)
#endif
#ifndef TPREFETCH
#define TPREFETCH 176         // tiles ahead whose staging data a CTA asks into L2 (0: off); 64 .. 300 measure the same
#endif

enum { KIND_COV = 0, KIND_EXT = 1, KIND_VAR = 2, KIND_VAR1 = 6 };    // (3 .. 5: the quality kinds below)

// Word arrays per kind (each word = two 16-bit counters, see the accessors below).
// NDIFF difference arrays come first (they are prefix-summed), then NPOINT point arrays.
template <int KIND> struct KindInfo;
template <> struct KindInfo<KIND_COV> { static const int NDIFF = 2, NPOINT = 0; };   // all[pp]: fwd | rev << 16
// strand s, class pair g = (all|pp), (mate unmapped|other chr), (same strand|faceaway), (softclipped|duplicate):
// word array 4 s + g, first class in the low half -- a typical read touches ONE word per interval end
template <> struct KindInfo<KIND_EXT> { static const int NDIFF = 8, NPOINT = 0; };
// all[pp], match[2 + pp], del[4 + pp] | ins[6 + pp], mismatches[8 + pp], mismatch class c = A,C,G,T,N: [10 + 2 c + pp]
template <> struct KindInfo<KIND_VAR> { static const int NDIFF = 6, NPOINT = 14; };
// variation without the strand split (opt.pyx:435-535 reports all / pp only): ONE word array per group, all reads in the
// low half, proper pairs in the high half -- all[0] match[1] del[2] | ins[3] mismatches[4] mismatch class c: [5 + c].
// Half the shared memory of KIND_VAR (four CTAs per SM), half the arrays to zero, scan and finalise.
template <> struct KindInfo<KIND_VAR1> { static const int NDIFF = 3, NPOINT = 7; };

// per-tile descriptor written by k_tile_desc (+ k_refpack for the flag)
struct __align__(16) TileDesc {
    int32_t x, y;        // candidate read index range
    int32_t n0;          // reads of the first chunk
    uint32_t flags;      // unused
    int64_t seq_lo;      // seq_off[x]
    int64_t seq_hi;      // seq_off[x + n0]
    int64_t cig_lo;      // CIGAR word span of the first chunk, cig_hi == cig_lo: not staged
    int64_t cig_hi;
};

static_assert(sizeof(TileDesc) == 48 && offsetof(TileDesc, x) == 0 && offsetof(TileDesc, y) == 4,
              "the tlen replay (k_pileup) reads a tile's candidate range as the int2 at the start of its descriptor");

struct TileArgs {
    const RInfo *ri;
    const uint32_t *cigar;
    const uint8_t *seq4;
    const int64_t *seq_off;
    const uint8_t *qual;       // quality kinds: qualities after the overlap rewrite
    const int64_t *qual_off;
    const uint32_t *lowq;      // min_baseq > 0: one bit per read base, set when its quality is below min_baseq (k_lowq)
    const TileDesc *desc;
    const uint32_t *refpack;   // REFN_WORDS words per tile (+ pad), KIND_VAR only
    const uint32_t *tflags;    // per tile, bit 0: a reference position of the tile has no 4-bit code (KIND_VAR)
    const uint8_t *ref;        // ASCII reference (slow path only)
    int64_t ref_start, ref_len;
    int64_t start, end;
    int64_t stride;            // elements between consecutive fields of `out`
    int32_t no_del;
    int32_t min_baseq;         // > 0 with lowq set: events of bases below it are dropped (opt.pyx:1900-1903)
    int32_t need_raw;          // read-level filters active: the row mask needs its own counter
    int32_t narrow;            // 1: out is uint16 [field][stride], 2: uint8 [field][stride] (padded rows): the host entry points' transport
    int32_t *ovf8;             // narrow == 2: set when a count does not fit 8 bits
    int32_t *out;
    uint8_t *rowmask;
    SmallState *st;
    unsigned long long *timing;   // TILE_TIMING builds only
};

// ------------------------------------------------------------------------- small helpers

__constant__ uint8_t c_refcode[256];   // upper-cased reference character -> 4-bit code, 16 = none

static void refcode_table(uint8_t *t) {
    memset(t, 16, 256);
    const char *nt16 = "=ACMGRSVTWYHKDBN";   // opt.pyx:72
    for (int k = 0; k < 16; ++k) {
        t[(unsigned char)nt16[k]] = (uint8_t)k;
        if (nt16[k] >= 'A' && nt16[k] <= 'Z') t[(unsigned char)(nt16[k] + 32)] = (uint8_t)k;   // refbase.upper()
    }
}

// 4-bit code -> mismatch class A,C,G,T,N,other (0..5)
__device__ __forceinline__ int mm_class(int code) {
    // nibble LUTs: codes 0-7 -> 5,0,1,5,2,5,5,5 ; codes 8-15 -> 3,5,5,5,5,5,5,4
    uint32_t lut = (code & 8) ? 0x45555553u : 0x55525105u;
    return (lut >> ((code & 7) << 2)) & 15;
}

struct TileCtx {
    uint32_t *cnt;          // counter arrays, [array][TT]
    const uint32_t *seqw;   // staged sequence bytes viewed as words
    const uint32_t *refn;   // reference tile, 4-bit codes, first position in the top nibble
    const uint8_t *ref;     // ASCII reference, for the slow path
    const uint8_t *qual;    // staged base qualities (quality kinds only)
    uint16_t *ev;           // this thread's deferred mismatch events, slot e at ev[e * EV_STRIDE] (LayVar)
    int64_t ref_start, ref_len;
    int t0, t1;             // tile position range (t1 = t0 + TT)
    int no_del;
    bool slow;              // a reference position without a code: compare base by base
};

// Counter words: bank = rev | pp << 1 lives in word array `group + (bank >> 1)`, forward strand
// in the low half, reverse in the high half.  Adding 1 << 16 or subtracting it (mod 2^32) keeps
// the word equal to hi * 65536 + lo as integers, so prefix sums of the words are the packed
// prefix sums as long as every final half is in [0, 65535].
enum { W_ALL = 0, W_MATCH = 2, W_DEL = 4, W_INS = 6, W_MMT = 8, W_MMC = 10 };
__device__ __forceinline__ uint32_t half_inc(int rev) { return 1u << ((rev & 1) << 4); }

__device__ __forceinline__ void interval_add(uint32_t *arr, int x, int l, int t0, uint32_t inc) {
    int a = max(x - t0, 0), b = min(x + l - t0, TT);
    if (a < b) {
        atomicAdd(arr + a, inc);
        if (b < TT) atomicAdd(arr + b, 0u - inc);
    }
}

// Mismatching bases are rare per thread (0.6 % of the bases) but frequent per warp: in four of five
// iterations of the word loop SOME lane has one, and handling it on the spot makes the whole warp
// step through ~30 instructions and 4-5 atomics for one or two active lanes.  A layout with
// DEFER set only records the event (tile position | 4-bit code << 9) in a per-thread slot of
// shared memory and applies it after the walk.  Measured on the 250 Mb workload: 6 % fewer warp
// instructions, no change in kernel time (13.50 vs 13.37 ms) -- the kernel is bound by the
// exposed latency between the phases of a tile, not by issue slots -- so it is off.
#define EV_SLOTS 4
#define EV_STRIDE TTHREADS

// ---- variation: one read per thread, one 8-base word per loop iteration ---------------------
//
// Counter arrays (per bank b = rev | pp << 1):
//   all[b]        difference array: reads covering the position (M, D and N spans)
//   match[4 + b]  difference array: +1 / -1 around every aligned span, and -1 / +1 around every
//                 mismatching base, so that its prefix sum IS the number of matches
//   del[8 + b]    difference array over D spans
//   ins[12 + b]   point: insertions (at the base before the I)
//   mmt[16 + b]   point: mismatches
//   mmc[20 + 4 c + b]  point: mismatching bases of class c = A, C, G, T, N
// A position's A count is mmc[A] + (reference is A ? matches : 0), etc..

// Counter layouts of the read walk below.  LayVar: the variation statistics, 16-bit counter pairs
// (bank >> 1 selects the word array, the strand the half).  LayBqx: baseq_ext(_strand), one
// 32-bit word array per bank (so any depth fits one pass; root mean squares are not additive
// over passes) and the squared qualities of mismatching bases as 64-bit point sums.
// LayVar1: variation without the strand split, one word array per group, all | proper pair << 16 (PER = word arrays per group).
// LayBqx2: baseq_ext without the strand split, TWO unpacked banks (pp or not): PER = 3 addresses like PER = 2 (bank >> 1)
// and counts like PER = 4 (full words).
struct LayVar { static const bool PACKED = true, VAR = true, DEFER = false; enum { PER = 2, ALL = W_ALL, MATCH = W_MATCH, DEL = W_DEL, INS = W_INS, MMT = W_MMT, MMC = W_MMC, SQMM = 0 }; };
struct LayVar1 { static const bool PACKED = true, VAR = true, DEFER = false; enum { PER = 1, ALL = 0, MATCH = 1, DEL = 2, INS = 3, MMT = 4, MMC = 5, SQMM = 0 }; };
struct LayBqx { static const bool PACKED = false, VAR = false, DEFER = false; enum { PER = 4, ALL = 0, MATCH = 4, DEL = 0, INS = 0, MMT = 8, MMC = 0, SQMM = 20 }; };
struct LayBqx2 { static const bool PACKED = false, VAR = false, DEFER = false; enum { PER = 3, ALL = 0, MATCH = 2, DEL = 0, INS = 0, MMT = 4, MMC = 0, SQMM = 10 }; };
template <class L> __device__ __forceinline__ uint32_t *carr(const TileCtx &T, int group, int bank) {
    return T.cnt + (group + ((L::PER == 2 || L::PER == 3) ? (bank >> 1) : L::PER == 1 ? 0 : bank)) * TT;
}
template <class L> __device__ __forceinline__ uint32_t cinc(int bank) {
    return L::PER == 2 ? half_inc(bank) : L::PER == 1 ? (1u | ((uint32_t)(bank >> 1) << 16)) : 1u;
}
template <int KIND> struct KindLay { typedef LayVar type; };
template <> struct KindLay<KIND_VAR1> { typedef LayVar1 type; };

// qidx: index of the base's quality in T.qual (LayBqx), < 0: no quality (reads without sequence)
template <class L>
__device__ __forceinline__ void var_mismatch(const TileCtx &T, int bank, int p, int code, int qidx) {
    const uint32_t inc = cinc<L>(bank);
    uint32_t *m = carr<L>(T, L::MATCH, bank) + p;
    atomicAdd(m, 0u - inc);
    if (p + 1 < TT) atomicAdd(m + 1, inc);
    atomicAdd(carr<L>(T, L::MMT, bank) + p, inc);
    if (L::VAR) {
        const int c = mm_class(code);
        if (c < 5) atomicAdd(carr<L>(T, L::MMC + L::PER * c, bank) + p, inc);
    } else {
        const unsigned long long q = qidx >= 0 ? T.qual[qidx] : 0;
        if (q) atomicAdd(reinterpret_cast<unsigned long long *>(T.cnt + L::SQMM * TT) + (L::PER == 3 ? (bank >> 1) : bank) * TT + p, q * q);
    }
}

struct VarCursor {        // per-thread walk state
    const uint32_t *cg;   // CIGAR words of the read (staged in shared memory, else global)
    int nc, k;            // number of ops, next op
    int x, y;             // reference / query cursor at op k
    int rn;               // tile-relative reference index of the next base of the current span
    int rem;              // bases left in the current span (inside the tile)
    uint32_t sn;          // nibble index of the next base in the staged sequence stream
    int qd;               // query index of a base = its tile-relative position + qd (current span)
    int w0, wstep;        // this thread owns the 8-base words w0/8, w0/8 + wstep/8, ... of every span
};

// Advance to the next aligned span that overlaps the tile, applying the side effects of the
// ops passed on the way (intervals of D spans, insertion points, the span's match interval).
template <class L>
__device__ __forceinline__ bool var_next_span(const TileCtx &T, VarCursor &c, int bank, uint32_t sn0, bool fx) {
    while (c.k < c.nc && c.x < T.t1) {
        const uint32_t w = c.cg[c.k];
        const int op = w & 15, l = (int)(w >> 4);
        const int k = c.k++;
        if (op_is_match(op)) {
            const int x = c.x, y = c.y;
            c.x += l; c.y += l;
            if (x + l > T.t0) {
                const int a = max(x - T.t0, 0), b = min(x + l - T.t0, TT);
                uint32_t *mt = carr<L>(T, L::MATCH, bank);
                const uint32_t inc = cinc<L>(bank);
                if (fx) {
                    atomicAdd(mt + a, inc);
                    if (b < TT) atomicAdd(mt + b, 0u - inc);
                    if (T.no_del) interval_add(carr<L>(T, L::ALL, bank), x, l, T.t0, inc);   // D / N spans do not count
                }
                // htslib indel > 0 at the last base of the op (A.1), opt.pyx:501
                const int li = x + l - 1 - T.t0;
                if (L::VAR && fx && k + 1 < c.nc && li >= 0 && li < TT) {
                    const uint32_t w2 = c.cg[k + 1];
                    const int op2 = w2 & 15;
                    bool ins = false;
                    if (op2 == OP_I) ins = (w2 >> 4) > 0;
                    else if (op2 == OP_P && k + 2 < c.nc) {
                        long long l3 = 0;
                        for (int j = k + 2; j < c.nc; ++j) {
                            const uint32_t w3 = c.cg[j];
                            const int op3 = w3 & 15;
                            if (op3 == OP_I) l3 += (w3 >> 4);
                            else if (op3 == OP_D || op3 == OP_M || op3 == OP_N || op3 == OP_EQ || op3 == OP_X) break;
                        }
                        ins = l3 > 0;
                    }
                    if (ins) atomicAdd(carr<L>(T, L::INS, bank) + li, inc);
                }
                c.qd = y + T.t0 - x;
                c.rn = a + c.w0; c.rem = b - a - c.w0;
                c.sn = sn0 + (uint32_t)(y + (a + T.t0 - x) + c.w0);
                if (c.rem > 0) return true;
                continue;                      // this part owns no word of the span
            }
        } else if (op == OP_D || op == OP_N) {
            if (L::VAR && fx && op == OP_D && !T.no_del && c.x + l > T.t0)
                interval_add(carr<L>(T, L::DEL, bank), c.x, l, T.t0, cinc<L>(bank));
            c.x += l;
        } else if (op == OP_I || op == OP_S) c.y += l;
    }
    return false;
}

// `part` of `nparts` threads share a read: part k compares the 8-base words k, k + nparts, ... of
// every aligned span; part 0 also applies the interval / point side effects.
// qb: index of the read's first base quality in T.qual (LayBqx only)
template <class L>
__device__ __forceinline__ void var_read(const TileCtx &T, const uint32_t *cg, const RInfo &r, uint32_t sn0, uint32_t *raw,
                                         int part, int nparts, int qb = 0) {
    const uint32_t cls = r.cls;
    if (!(cls & CLS_ADMITTED)) return;
    const bool fx = part == 0;
    if (raw && fx) interval_add(raw, r.pos, r.end - r.pos, T.t0, 1u);
    if (!(cls & CLS_COUNTED)) return;
    const int bank = cls & 3u;
    const uint32_t inc = cinc<L>(bank);
    if (!T.no_del && fx) interval_add(carr<L>(T, L::ALL, bank), r.pos, r.end - r.pos, T.t0, inc);   // M, D and N all count
    if (cls & CLS_NOREF) {   // mapped, CIGAR consumes no reference: one deletion event at pos
        if (L::VAR && !T.no_del && fx) interval_add(carr<L>(T, L::DEL, bank), r.pos, 1, T.t0, inc);
        return;
    }
    VarCursor c;
    c.cg = cg; c.nc = (int)r.n_cigar; c.k = 0; c.x = r.pos; c.y = 0;
    c.w0 = part << 3; c.wstep = nparts << 3;
    const bool simple = cls & CLS_SIMPLE;
    bool have;
    if (simple) {
        const int a = max(r.pos - T.t0, 0), b = min(r.end - T.t0, TT);
        if (a < b && fx) {
            uint32_t *mt = carr<L>(T, L::MATCH, bank);
            atomicAdd(mt + a, inc);
            if (b < TT) atomicAdd(mt + b, 0u - inc);
            if (T.no_del) interval_add(carr<L>(T, L::ALL, bank), r.pos, r.end - r.pos, T.t0, inc);
        }
        c.qd = T.t0 - r.pos;
        c.rn = a + c.w0; c.rem = b - a - c.w0; c.sn = sn0 + (uint32_t)(a + T.t0 - r.pos + c.w0);
        have = c.rem > 0;
    } else have = var_next_span<L>(T, c, bank, sn0, fx);
    const bool hasseq = cls & CLS_HASSEQ;
    if (!hasseq || T.slow) {
        while (have) {
            const int n8 = min(c.rem, 8);
            if (!hasseq) {       // no sequence: every base is a mismatch of no class (opt.pyx:2146-2147)
                for (int i = 0; i < n8; ++i) var_mismatch<L>(T, bank, c.rn + i, 0, -1);
            } else {             // a reference position of this tile has no 4-bit code
                for (int i = 0; i < n8; ++i) {
                    const uint32_t k = c.sn + i;
                    const int code = (T.seqw[k >> 3] >> (((k & 7) ^ 1) << 2)) & 15;   // byte k / 2, high nibble first
                    const int64_t ro = (int64_t)T.t0 + c.rn + i - T.ref_start;
                    const int rc = (T.ref && ro >= 0 && ro < T.ref_len) ? c_refcode[__ldg(T.ref + ro)] : 16;
                    if (code != rc) var_mismatch<L>(T, bank, c.rn + i, code, qb + c.rn + i + c.qd);
                }
            }
            c.rn += c.wstep; c.sn += c.wstep; c.rem -= c.wstep;
            if (c.rem <= 0) have = simple ? false : var_next_span<L>(T, c, bank, sn0, fx);
        }
        return;
    }
    // fast path.  Within a span the word cursors advance by whole words (wstep = 8 nparts nibbles),
    // so the nibble shifts are per-span constants and the loop is: 4 loads, 2 byte swaps, 2 funnel
    // shifts, XOR.  The words of the NEXT iteration are loaded before the (rare, atomic) mismatch
    // handling of the current one, so the shared-memory latency overlaps it.
    const int wwords = c.wstep >> 3;
    int nev = 0;
    const uint32_t *sp = nullptr, *rp = nullptr;
    uint32_t ssh = 0, rsh = 0, s0 = 0, s1 = 0, r0 = 0, r1 = 0;
#define VAR_SPAN_BEGIN()                                                        \
    do {                                                                        \
        sp = T.seqw + (c.sn >> 3); ssh = (c.sn & 7) << 2;                         \
        rp = T.refn + ((uint32_t)c.rn >> 3); rsh = ((uint32_t)c.rn & 7) << 2;     \
        s0 = sp[0]; s1 = sp[1]; r0 = rp[0]; r1 = rp[1];                           \
    } while (0)
    if (have) VAR_SPAN_BEGIN();
    while (have) {
        const uint32_t sw = __funnelshift_l(__byte_perm(s1, 0, 0x0123), __byte_perm(s0, 0, 0x0123), ssh);
        uint32_t xr = sw ^ __funnelshift_l(r1, r0, rsh);
        if (c.rem < 8) xr &= 0xffffffffu << ((8 - c.rem) << 2);
        const int p0 = c.rn;
        const int qd0 = c.qd;
        c.rn += c.wstep; c.rem -= c.wstep;
        const bool in_span = c.rem > 0;
        if (in_span) {
            sp += wwords; rp += wwords;
            s0 = sp[0]; s1 = sp[1]; r0 = rp[0]; r1 = rp[1];
        }
        while (xr) {
            const int lz = __clz(xr) >> 2, sh = 28 - (lz << 2);
            if (L::DEFER && nev < EV_SLOTS) T.ev[nev++ * EV_STRIDE] = (uint16_t)((p0 + lz) | (((sw >> sh) & 15) << 9));
            else var_mismatch<L>(T, bank, p0 + lz, (sw >> sh) & 15, qb + p0 + lz + qd0);
            xr &= ~(0xfu << sh);
        }
        if (!in_span) {
            have = simple ? false : var_next_span<L>(T, c, bank, sn0, fx);
            if (have) VAR_SPAN_BEGIN();
        }
    }
#undef VAR_SPAN_BEGIN
    if (L::DEFER) {
#pragma unroll 1
        for (int e = 0; e < nev; ++e) {
            const uint32_t w = T.ev[e * EV_STRIDE];
            var_mismatch<L>(T, bank, (int)(w & 511u), (int)(w >> 9), -1);
        }
    }
}

// does an insertion follow CIGAR op k (htslib `indel > 0` at the last base of the op, A.1)?
__device__ __forceinline__ bool ins_follows(const uint32_t *cg, int nc, int k) {
    if (k + 1 >= nc) return false;
    const uint32_t w2 = cg[k + 1];
    const int op2 = w2 & 15;
    if (op2 == OP_I) return (w2 >> 4) > 0;
    if (op2 == OP_P && k + 2 < nc) {
        long long l3 = 0;
        for (int j = k + 2; j < nc; ++j) {
            const uint32_t w3 = cg[j];
            const int op3 = w3 & 15;
            if (op3 == OP_I) l3 += (w3 >> 4);
            else if (op3 == OP_D || op3 == OP_M || op3 == OP_N || op3 == OP_EQ || op3 == OP_X) break;
        }
        return l3 > 0;
    }
    return false;
}

// ---- variation, balanced walk -----------------------------------------------------------------
//
// var_read gives a read to one or two threads, so a thread's work is the read's overlap with the tile
// (0 .. 19 words) and the threads beyond 2 n idle; everybody then waits at the barrier for the longest.
// The balanced walk splits the chunk in two steps: (A) one thread per read applies what is constant
// along spans (interval adds, insertion points) and emits a record per aligned span that overlaps the
// tile; (B) after a prefix sum over the spans' word counts every thread takes the same number of
// consecutive 8-base words, wherever they fall among the spans.
// Measured on the 250 Mb workload (parity-green): 14.7 ms against 13.6 ms for var_read -- the extra block
// barrier, the prefix sums and the per-thread binary search cost more than the balance gains -- so it is off.
#ifndef TBAL
#define TBAL 0               // 1: balanced walk; 0: the read-per-thread walk (var_read) for every chunk
#endif
#define SPAN_MAX (TCH + 64)  // span records per chunk

struct SpanRec;
__device__ __forceinline__ void var_walk_span(const TileCtx &T, const SpanRec &sr, int k, int kend);

struct __align__(16) SpanRec {
    uint32_t a_len;      // tile-relative start | length << 16 (length <= TT)
    uint32_t sn;         // nibble index of the span's first base in the staged sequence stream
    uint32_t bank;
    uint32_t wbase;      // exclusive prefix sum of the spans' word counts
};

// step A for one read (spans beyond the list's capacity are walked on the spot)
__device__ __forceinline__ bool var_emit_spans(const TileCtx &T, const uint32_t *cg, const RInfo &r, uint32_t sn0, uint32_t *raw,
                                               SpanRec *spans, int *nspan) {
    const uint32_t cls = r.cls;
    if (!(cls & CLS_ADMITTED)) return true;
    if (raw) interval_add(raw, r.pos, r.end - r.pos, T.t0, 1u);
    if (!(cls & CLS_COUNTED)) return true;
    const int bank = cls & 3u;
    const uint32_t inc = half_inc(bank);
    if (!T.no_del) interval_add(T.cnt + (W_ALL + (bank >> 1)) * TT, r.pos, r.end - r.pos, T.t0, inc);   // M, D and N all count
    if (cls & CLS_NOREF) {   // mapped, CIGAR consumes no reference: one deletion event at pos
        if (!T.no_del) interval_add(T.cnt + (W_DEL + (bank >> 1)) * TT, r.pos, 1, T.t0, inc);
        return true;
    }
    uint32_t *mt = T.cnt + (W_MATCH + (bank >> 1)) * TT;
    auto span = [&](int x, int l, int y) -> bool {
        const int a = max(x - T.t0, 0), b = min(x + l - T.t0, TT);
        if (a >= b) return true;
        atomicAdd(mt + a, inc);
        if (b < TT) atomicAdd(mt + b, 0u - inc);
        if (T.no_del) interval_add(T.cnt + (W_ALL + (bank >> 1)) * TT, x, l, T.t0, inc);   // D / N spans do not count
        SpanRec s;
        s.a_len = (uint32_t)a | ((uint32_t)(b - a) << 16);
        s.sn = sn0 + (uint32_t)(y + (a + T.t0 - x));
        s.bank = (uint32_t)bank;
        s.wbase = 0;
        const int k = atomicAdd(nspan, 1);
        if (k < SPAN_MAX) spans[k] = s;
        else var_walk_span(T, s, 0, (b - a + 7) >> 3);      // the list is full: this thread walks the span itself
        return true;
    };
    if (cls & CLS_SIMPLE) return span(r.pos, r.end - r.pos, 0);
    const int nc = (int)r.n_cigar;
    int x = r.pos, y = 0;
    for (int k = 0; k < nc && x < T.t1; ++k) {
        const uint32_t w = cg[k];
        const int op = w & 15, l = (int)(w >> 4);
        if (op_is_match(op)) {
            if (x + l > T.t0) {
                if (!span(x, l, y)) return false;
                const int li = x + l - 1 - T.t0;            // htslib indel > 0 at the last base of the op (A.1), opt.pyx:501
                if (li >= 0 && li < TT && ins_follows(cg, nc, k)) atomicAdd(T.cnt + (W_INS + (bank >> 1)) * TT + li, inc);
            }
            x += l; y += l;
        } else if (op == OP_D || op == OP_N) {
            if (op == OP_D && !T.no_del && x + l > T.t0) interval_add(T.cnt + (W_DEL + (bank >> 1)) * TT, x, l, T.t0, inc);
            x += l;
        } else if (op == OP_I || op == OP_S) y += l;
    }
    return true;
}

// words [k, kend) of one span
__device__ __forceinline__ void var_walk_span(const TileCtx &T, const SpanRec &sr, int k, int kend) {
    const int a = (int)(sr.a_len & 0xffffu), len = (int)(sr.a_len >> 16);
    const int bank = (int)sr.bank;
    const uint32_t sn = sr.sn + 8u * (uint32_t)k;
    int p0 = a + 8 * k;
    const uint32_t *sp = T.seqw + (sn >> 3), *rp = T.refn + ((uint32_t)p0 >> 3);
    const uint32_t ssh = (sn & 7u) << 2, rsh = ((uint32_t)p0 & 7u) << 2;
    uint32_t s0 = __byte_perm(sp[0], 0, 0x0123), r0 = rp[0];
    for (; k < kend; ++k, p0 += 8) {
        ++sp; ++rp;
        const uint32_t s1 = __byte_perm(sp[0], 0, 0x0123), r1 = rp[0];
        const uint32_t sw = __funnelshift_l(s1, s0, ssh);
        uint32_t xr = sw ^ __funnelshift_l(r1, r0, rsh);
        const int rem = len - 8 * k;
        if (rem < 8) xr &= 0xffffffffu << ((8 - rem) << 2);
        s0 = s1; r0 = r1;
        while (xr) {
            const int lz = __clz(xr) >> 2, sh = 28 - (lz << 2);
            var_mismatch<LayVar>(T, bank, p0 + lz, (sw >> sh) & 15, -1);
            xr &= ~(0xfu << sh);
        }
    }
}

// step B: words [g0, g1) of the chunk's span list
__device__ __forceinline__ void var_walk_words(const TileCtx &T, const SpanRec *spans, int nspan, int g0, int g1) {
    if (g0 >= g1) return;
    int lo = 0, hi = nspan - 1;                      // last span whose wbase <= g0
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int)spans[mid].wbase <= g0) lo = mid; else hi = mid - 1; }
    int i = lo, g = g0;
    while (g < g1) {
        const SpanRec sr = spans[i];
        const int nw = (int)((sr.a_len >> 16) + 7) >> 3;
        const int k = g - (int)sr.wbase;
        const int kend = min(nw, k + (g1 - g));
        var_walk_span(T, sr, k, kend);
        g += kend - k;
        ++i;
    }
}

// ---- coverage kinds: one read per thread ------------------------------------------------------

template <int KIND>
__device__ __forceinline__ void cover_segment(const TileCtx &T, uint32_t cls, int x, int l) {
    const uint32_t inc = half_inc(cls & CLS_REV ? 1 : 0);
    if (KIND == KIND_COV) interval_add(T.cnt + ((cls >> 1) & 1u) * TT, x, l, T.t0, inc);
    else {   // KIND_EXT: word arrays of class pairs, per strand
        const uint32_t m = (cls >> CLS_EXT_SHIFT) & 255u;
        const int a = max(x - T.t0, 0), b = min(x + l - T.t0, TT);
        if (a >= b) return;
        uint32_t *base = T.cnt + (cls & CLS_REV ? 4 : 0) * TT;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint32_t bits = (m >> (2 * g)) & 3u;
            if (bits) {
                const uint32_t v = (bits & 1u) | ((bits >> 1) << 16);
                atomicAdd(base + g * TT + a, v);
                if (b < TT) atomicAdd(base + g * TT + b, 0u - v);
            }
        }
    }
}

template <int KIND>
__device__ void cover_read(const TileCtx &T, const TileArgs &A, const RInfo &r, uint32_t *raw) {
    const uint32_t cls = r.cls;
    if (!(cls & CLS_ADMITTED)) return;
    if (raw) interval_add(raw, r.pos, r.end - r.pos, T.t0, 1u);
    if (!(cls & CLS_COUNTED)) return;
    if (!T.no_del || (cls & CLS_SIMPLE)) {   // D / N / zero-length alignments all count (opt.pyx:1904-1906)
        cover_segment<KIND>(T, cls, r.pos, r.end - r.pos);
        return;
    }
    if (cls & CLS_NOREF) return;             // its only event is deletion-like: dropped by no_del
    const uint32_t *cg = A.cigar + r.cigar_off;
    const int nc = (int)r.n_cigar;
    int x = r.pos;
    for (int k = 0; k < nc && x < T.t1; ++k) {
        uint32_t w = __ldg(cg + k);
        int op = w & 15, l = (int)(w >> 4);
        if (op_is_match(op)) {
            if (x + l > T.t0) cover_segment<KIND>(T, cls, x, l);
            x += l;
        } else if (op == OP_D || op == OP_N) x += l;
    }
}

// ---- min_baseq > 0: corrections for the bases below the threshold -----------------------------
//
// stat_pileup drops the event of a read at a column when min_baseq > 0, the read has a base there
// (not a deletion) and its quality is below the threshold (opt.pyx:1900-1903).  The scatter above
// runs as if every base passed; k_lowq has marked the failing bases (one bit per base, word index
// (seq_off >> 4) + read index + qpos / 32), and the read's thread now takes each of them back out:
// -1 at the position in the read-count difference arrays (+1 at the next), and for the variation
// kinds the match / mismatch / base-class / insertion contribution of that base.  Everything is
// linear in the counters, so the order against the main pass does not matter.

__device__ __forceinline__ void lowq_prefetch(const TileArgs &A, const RInfo &r, int ridx) {
    // the read's mark words are wanted after its walk: start them towards L1 now (a 150 bp read: 20 bytes)
    const uint32_t *mw = A.lowq + (r.seq_off >> 4) + ridx;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(mw));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(mw + 4));
}

// calls f(p, qi, last) for every marked base of the aligned span [x, x + l) (first base query index y)
// that lies in the tile: p tile-relative position, qi query index, last: it is the last base of the span
template <class F>
__device__ __forceinline__ void lowq_span(const uint32_t *mw, int x, int l, int y, int t0, F f, int part = 0, int nparts = 1) {
    const int a = max(x - t0, 0), b = min(x + l - t0, TT);
    if (a >= b) return;
    const int qa = y + (a + t0 - x), qb = y + (b + t0 - x);       // query range [qa, qb)
    const int w0 = qa >> 5, w1 = (qb - 1) >> 5;
    auto word = [&](uint32_t m, int w) {
        const int base = w << 5;
        if (base < qa) m &= 0xffffffffu << (qa - base);
        if (base + 32 > qb) m &= 0xffffffffu >> (base + 32 - qb);
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int qi = base + bit;
            f(a + (qi - qa), qi, qi == y + l - 1);
        }
    };
    for (int w = w0 + part; w <= w1; w += nparts) word(__ldg(mw + w), w);       // the threads sharing a read take words in turn
}

// The same over maximal RUNS of marked bases: calls f(p, len) with the tile-relative position of a run's first base.
// What a dropped event takes out of an interval-add counter is linear, so a run costs one interval (two atomics)
// whatever its length -- and marks come in runs: the mate-overlap rewrite zeroes the whole overlap of one mate
// (SURVEY.md A.3), 50 bases and more in a row.  (A run that crosses a 32-bit mark word is reported in two pieces.)
template <class F>
__device__ __forceinline__ void lowq_runs(const uint32_t *mw, int x, int l, int y, int t0, F f) {
    const int a = max(x - t0, 0), b = min(x + l - t0, TT);
    if (a >= b) return;
    const int qa = y + (a + t0 - x), qb = y + (b + t0 - x);       // query range [qa, qb)
    const int w0 = qa >> 5, w1 = (qb - 1) >> 5;
    for (int w = w0; w <= w1; ++w) {
        uint32_t m = __ldg(mw + w);
        const int base = w << 5;
        if (base < qa) m &= 0xffffffffu << (qa - base);
        if (base + 32 > qb) m &= 0xffffffffu >> (base + 32 - qb);
        while (m) {
            const int bit = __ffs(m) - 1;
            const uint32_t inv = ~(m >> bit);
            const int len = inv ? __ffs(inv) - 1 : 32 - bit;
            f(a + (base + bit - qa), len);
            m = (len + bit >= 32) ? 0u : (m & (0xffffffffu << (bit + len)));
        }
    }
}

// variation kinds (LayVar).  Reads without sequence have quality 0 at every base: they keep only their
// deletion / ref-skip events, which var_read is told through `drop_bases`.
template <class L>
__device__ void var_lowq_fix(const TileCtx &T, const TileArgs &A, const uint32_t *cg, const RInfo &r, uint32_t sn0, int ridx,
                             int part, int nparts) {
    const uint32_t cls = r.cls;
    if (part != 0 || !(cls & CLS_COUNTED) || (cls & CLS_NOREF) || !(cls & CLS_HASSEQ)) return;
    const int bank = cls & 3u;
    const uint32_t inc = cinc<L>(bank);
    const uint32_t *mw = A.lowq + (r.seq_off >> 4) + ridx;
    const int nc = (int)r.n_cigar;
    int x = r.pos, y = 0;
    for (int k = 0; k < nc && x < T.t1; ++k) {
        const uint32_t w = cg[k];
        const int op = w & 15, l = (int)(w >> 4);
        if (op_is_match(op)) {
            if (x + l > T.t0) {
                const bool ins = ins_follows(cg, nc, k);
                // a RUN of marked bases [p, p + len) leaves the read count and the match count as one interval each (the
                // mate-overlap rewrite marks whole overlaps); only its mismatching bases -- which the walk had already
                // taken out of the match count -- are handled one by one
                lowq_runs(mw, x, l, y, T.t0, [&](int p, int len) {
                    const int e = p + len;
                    uint32_t *al = carr<L>(T, L::ALL, bank), *mt = carr<L>(T, L::MATCH, bank);
                    atomicAdd(al + p, 0u - inc);
                    atomicAdd(mt + p, 0u - inc);
                    if (e < TT) { atomicAdd(al + e, inc); atomicAdd(mt + e, inc); }
                    const int qi0 = y + (T.t0 + p - x);
                    auto mismatch = [&](int pj, int code) {
                        atomicAdd(mt + pj, inc);                       // it was no match: undo the interval here
                        if (pj + 1 < TT) atomicAdd(mt + pj + 1, 0u - inc);
                        atomicAdd(carr<L>(T, L::MMT, bank) + pj, 0u - inc);
                        const int c = mm_class(code);
                        if (c < 5) atomicAdd(carr<L>(T, L::MMC + L::PER * c, bank) + pj, 0u - inc);
                    };
                    if (T.slow) {               // a reference position of this tile has no 4-bit code: base by base
                        for (int j = 0; j < len; ++j) {
                            const uint32_t kk = sn0 + (uint32_t)(qi0 + j);
                            const int code = (T.seqw[kk >> 3] >> (((kk & 7) ^ 1) << 2)) & 15;
                            const int64_t ro = (int64_t)T.t0 + p + j - T.ref_start;
                            const int rc = (T.ref && ro >= 0 && ro < T.ref_len) ? c_refcode[__ldg(T.ref + ro)] : 16;
                            if (code != rc) mismatch(p + j, code);
                        }
                    } else {                    // eight bases per step, as the walk compares them
                        for (int j = 0; j < len; j += 8) {
                            const uint32_t sn = sn0 + (uint32_t)(qi0 + j), rn = (uint32_t)(p + j);
                            const uint32_t *sp = T.seqw + (sn >> 3), *rp = T.refn + (rn >> 3);
                            const uint32_t sw = __funnelshift_l(__byte_perm(sp[1], 0, 0x0123), __byte_perm(sp[0], 0, 0x0123), (sn & 7) << 2);
                            uint32_t xr = sw ^ __funnelshift_l(rp[1], rp[0], (rn & 7) << 2);
                            if (len - j < 8) xr &= 0xffffffffu << ((8 - (len - j)) << 2);
                            while (xr) {
                                const int lz = __clz(xr) >> 2, sh = 28 - (lz << 2);
                                mismatch(p + j + lz, (sw >> sh) & 15);
                                xr &= ~(0xfu << sh);
                            }
                        }
                    }
                    if (ins && qi0 + len - 1 == y + l - 1) atomicAdd(carr<L>(T, L::INS, bank) + e - 1, 0u - inc);
                });
            }
            x += l; y += l;
        } else if (op == OP_D || op == OP_N) x += l;
        else if (op == OP_I || op == OP_S) y += l;
    }
}

// coverage kinds: the read-count words of every marked base
template <int KIND>
__device__ void cover_lowq_fix(const TileCtx &T, const TileArgs &A, const RInfo &r, int ridx) {
    const uint32_t cls = r.cls;
    if (!(cls & CLS_COUNTED) || (cls & CLS_NOREF) || !(cls & CLS_HASSEQ)) return;
    const uint32_t *mw = A.lowq + (r.seq_off >> 4) + ridx;
    const uint32_t *cg = A.cigar + r.cigar_off;
    const int nc = (int)r.n_cigar;
    auto undo = [&](int p, int len) {              // a run of marked bases [p, p + len): one interval per counter word
        const int e = p + len;
        if (KIND == KIND_COV) {
            const uint32_t inc = half_inc(cls & CLS_REV ? 1 : 0);
            uint32_t *al = T.cnt + ((cls >> 1) & 1u) * TT;
            atomicAdd(al + p, 0u - inc);
            if (e < TT) atomicAdd(al + e, inc);
        } else {
            const uint32_t m = (cls >> CLS_EXT_SHIFT) & 255u;
            uint32_t *base = T.cnt + (cls & CLS_REV ? 4 : 0) * TT;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t bits = (m >> (2 * g)) & 3u;
                if (bits) {
                    const uint32_t v = (bits & 1u) | ((bits >> 1) << 16);
                    atomicAdd(base + g * TT + p, 0u - v);
                    if (e < TT) atomicAdd(base + g * TT + e, v);
                }
            }
        }
    };
    if (cls & CLS_SIMPLE) { lowq_runs(mw, r.pos, r.end - r.pos, 0, T.t0, undo); return; }
    int x = r.pos, y = 0;
    for (int k = 0; k < nc && x < T.t1; ++k) {
        const uint32_t w = __ldg(cg + k);
        const int op = w & 15, l = (int)(w >> 4);
        if (op_is_match(op)) {
            if (x + l > T.t0) lowq_runs(mw, x, l, y, T.t0, undo);
            x += l; y += l;
        } else if (op == OP_D || op == OP_N) x += l;
        else if (op == OP_I || op == OP_S) y += l;
    }
}

// a read without sequence under min_baseq > 0: only its deletion / ref-skip events remain
template <int KIND>
__device__ void noseq_minq_read(const TileCtx &T, const TileArgs &A, const RInfo &r, uint32_t *raw) {
    const uint32_t cls = r.cls;
    if (!(cls & CLS_ADMITTED)) return;
    if (raw) interval_add(raw, r.pos, r.end - r.pos, T.t0, 1u);
    if (!(cls & CLS_COUNTED) || T.no_del) return;
    const int bank = cls & 3u;
    if (cls & CLS_NOREF) {         // one deletion-like event at pos
        if (KIND == KIND_VAR || KIND == KIND_VAR1) {
            typedef typename KindLay<KIND>::type L;
            interval_add(carr<L>(T, L::ALL, bank), r.pos, 1, T.t0, cinc<L>(bank));
            interval_add(carr<L>(T, L::DEL, bank), r.pos, 1, T.t0, cinc<L>(bank));
        } else cover_segment<KIND>(T, cls, r.pos, 1);
        return;
    }
    const uint32_t *cg = A.cigar + r.cigar_off;
    const int nc = (int)r.n_cigar;
    int x = r.pos;
    for (int k = 0; k < nc && x < T.t1; ++k) {
        const uint32_t w = __ldg(cg + k);
        const int op = w & 15, l = (int)(w >> 4);
        if (op_is_match(op)) x += l;
        else if (op == OP_D || op == OP_N) {
            if (x + l > T.t0) {
                if (KIND == KIND_VAR || KIND == KIND_VAR1) {
                    typedef typename KindLay<KIND>::type L;
                    interval_add(carr<L>(T, L::ALL, bank), x, l, T.t0, cinc<L>(bank));
                    if (op == OP_D) interval_add(carr<L>(T, L::DEL, bank), x, l, T.t0, cinc<L>(bank));
                } else cover_segment<KIND>(T, cls, x, l);
            }
            x += l;
        }
    }
}

// ---- in-place inclusive prefix sum of one N-word array (TT by default) by one warp -----------
// lane l owns words [128 k + 4 l, 128 k + 4 l + 4) for k = 0 .. N/128-1 (conflict-free 128-bit
// accesses).
template <int N = TT>
__device__ __forceinline__ void warp_scan_array(uint32_t *arr, int lane) {
    const int NK = N / 128;
    uint4 v[NK];
    uint32_t s[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        v[k] = *reinterpret_cast<uint4 *>(arr + 128 * k + 4 * lane);
        v[k].y += v[k].x; v[k].z += v[k].y; v[k].w += v[k].z;
        s[k] = v[k].w;
    }
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        uint32_t inc = s[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        uint32_t off = carry + inc - s[k];
        carry += __shfl_sync(FULL, inc, 31);
        v[k].x += off; v[k].y += off; v[k].z += off; v[k].w += off;
        *reinterpret_cast<uint4 *>(arr + 128 * k + 4 * lane) = v[k];
    }
}

// ---- finalisation: rec() of the reference classes, all linear in the counters ------------------

__device__ __forceinline__ uint4 operator+(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint4 lo16(uint4 a) { return make_uint4(a.x & 0xffffu, a.y & 0xffffu, a.z & 0xffffu, a.w & 0xffffu); }
__device__ __forceinline__ uint4 hi16(uint4 a) { return make_uint4(a.x >> 16, a.y >> 16, a.z >> 16, a.w >> 16); }

enum { OUT_STORE = 0, OUT_ACCUM = 1, OUT_NARROW = 2, OUT_NARROW8 = 3 };   // int32 store / int32 add / 16-bit store / 8-bit store

template <int MODE>
struct QuadOut {
    // position 4q of the next field to write; int32 elements, or uint16 / uint8 elements in the
    // OUT_NARROW / OUT_NARROW8 modes (then p advances in those units: the pointer is reinterpreted)
    int32_t *p;
    int64_t stride;
    bool vec;          // vector stores allowed
    int nvalid;        // positions of the quad inside [start, end)
    uint32_t seen;     // OUT_NARROW8: OR of every value written (a bit >= 8 set: the chunk needs 16 bits)
    __device__ __forceinline__ void put(uint4 v) {
        if (MODE == OUT_NARROW8) {     // rows are padded and 4-byte aligned in this mode
            uint8_t *h = reinterpret_cast<uint8_t *>(p);
            seen |= v.x | v.y | v.z | v.w;
            *reinterpret_cast<uint32_t *>(h) = (v.x & 255u) | ((v.y & 255u) << 8) | ((v.z & 255u) << 16) | (v.w << 24);
            p = reinterpret_cast<int32_t *>(h + stride);
            return;
        }
        if (MODE == OUT_NARROW) {      // rows are padded and 8-byte aligned in this mode
            uint16_t *h = reinterpret_cast<uint16_t *>(p);
            *reinterpret_cast<uint2 *>(h) = make_uint2(v.x | (v.y << 16), v.z | (v.w << 16));
            p = reinterpret_cast<int32_t *>(h + stride);
            return;
        }
        if (vec) {
            if (MODE == OUT_ACCUM) v = v + *reinterpret_cast<const uint4 *>(p);
#if TSTCS
            __stcs(reinterpret_cast<uint4 *>(p), v);       // written once, never read again: streaming store
#else
            *reinterpret_cast<uint4 *>(p) = v;
#endif
        } else {
            if (nvalid > 0) p[0] = (MODE == OUT_ACCUM ? p[0] : 0) + v.x;
            if (nvalid > 1) p[1] = (MODE == OUT_ACCUM ? p[1] : 0) + v.y;
            if (nvalid > 2) p[2] = (MODE == OUT_ACCUM ? p[2] : 0) + v.z;
            if (nvalid > 3) p[3] = (MODE == OUT_ACCUM ? p[3] : 0) + v.w;
        }
        p += stride;
    }
    __device__ __forceinline__ void skip(int nfields) {
        if (MODE == OUT_NARROW8) p = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(p) + stride * nfields);
        else if (MODE == OUT_NARROW) p = reinterpret_cast<int32_t *>(reinterpret_cast<uint16_t *>(p) + stride * nfields);
        else p += stride * nfields;
    }
    // w0 = not-proper-pair word (fwd | rev << 16), w1 = proper-pair word ->
    // all, fwd, rev, pp, pp_fwd, pp_rev (CountPpStrand, opt.pyx:147-168) or all, pp
    template <bool STRAND>
    __device__ __forceinline__ void put_group(uint4 w0, uint4 w1) {
        const uint4 s = w0 + w1;                 // halves stay below 65536: fwd | rev
        const uint4 fwd = lo16(s), rev = hi16(s), pf = lo16(w1), pr = hi16(w1);
        put(fwd + rev);
        if (STRAND) { put(fwd); put(rev); put(pf + pr); put(pf); put(pr); }
        else put(pf + pr);
    }
};

__device__ __forceinline__ uint4 ld4(const uint32_t *cnt, int arr, int q) {
    return *reinterpret_cast<const uint4 *>(cnt + arr * TT + 4 * q);
}
// m + (refcode of the position == code ? mt : 0) for the four positions of a quad (packed words)
__device__ __forceinline__ uint4 base_count(uint4 m, uint4 mt, uint32_t rc4, uint32_t code) {
    uint4 r = m;
    if (((rc4 >> 12) & 15u) == code) r.x += mt.x;
    if (((rc4 >> 8) & 15u) == code) r.y += mt.y;
    if (((rc4 >> 4) & 15u) == code) r.z += mt.z;
    if ((rc4 & 15u) == code) r.w += mt.w;
    return r;
}

// Writes the fields owned by (quad q, sub-part `part` of 4) and returns the quad's total read
// count when this part computes it (part 0), for the row mask.
// sub-parts a quad's fields are split into (tasks of tile_output): the per-task overhead (pointer arithmetic, range
// checks) is worth sharing between more fields when a statistic has few
template <int STAT> struct FinParts {
    static const int N = (STAT == PSS_COVERAGE || STAT == PSS_COVERAGE_STRAND || STAT == PSS_COVERAGE_GC) ? 1
                         : 4;      // (coverage_ext with 1 or 2: slower, the 192 threads are no longer all busy)
};

template <int STAT, int MODE>
__device__ __forceinline__ uint4 tile_finalize(const uint32_t *cnt, const uint32_t *refn, int q, int part,
                                               QuadOut<MODE> &O) {
    uint4 all = make_uint4(0, 0, 0, 0);
    if (STAT == PSS_COVERAGE || STAT == PSS_COVERAGE_STRAND || STAT == PSS_COVERAGE_GC) {
        if (part == 0) {
            const uint4 w0 = ld4(cnt, 0, q), w1 = ld4(cnt, 1, q);
            const uint4 s = w0 + w1;
            all = lo16(s) + hi16(s);
            if (STAT == PSS_COVERAGE_GC) O.skip(1);                      // field 0 = gc (k_gc)
            if (STAT == PSS_COVERAGE_STRAND) O.template put_group<true>(w0, w1);
            else O.template put_group<false>(w0, w1);
        }
    } else if (STAT == PSS_COVERAGE_EXT || STAT == PSS_COVERAGE_EXT_STRAND) {
        // word array g (forward) / 4 + g (reverse) holds classes 2 g (low half), 2 g + 1; a sub-part takes 4 / N arrays
        constexpr int GP = 4 / FinParts<STAT>::N;
        O.skip(part * GP * 2 * (STAT == PSS_COVERAGE_EXT ? 1 : 3));
#pragma unroll
        for (int gg = 0; gg < GP; ++gg) {
            const int g = part * GP + gg;
            const uint4 wf = ld4(cnt, g, q), wr = ld4(cnt, 4 + g, q);
            if (g == 0) all = lo16(wf) + lo16(wr);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint4 f = k ? hi16(wf) : lo16(wf), r = k ? hi16(wr) : lo16(wr);
                O.put(f + r);
                if (STAT == PSS_COVERAGE_EXT_STRAND) { O.put(f); O.put(r); }
            }
        }
    } else if (STAT == PSS_VARIATION) {
        // LayVar1: one word per group and position, all | proper pair << 16.  Groups and sub-parts as below.
        typedef LayVar1 L1;
        const int g0 = part == 0 ? 0 : part == 1 ? 3 : part == 2 ? 6 : 8, g1 = part == 0 ? 3 : part == 1 ? 6 : part == 2 ? 8 : 10;
        // counter array (4 bits) per group: reads, matches, mismatches, deletions, insertions, A, C, T, G, N
        const unsigned long long ARR = (unsigned long long)L1::ALL | ((unsigned long long)L1::MATCH << 4) | ((unsigned long long)L1::MMT << 8) |
                                       ((unsigned long long)L1::DEL << 12) | ((unsigned long long)L1::INS << 16) | ((unsigned long long)(L1::MMC + 0) << 20) |
                                       ((unsigned long long)(L1::MMC + 1) << 24) | ((unsigned long long)(L1::MMC + 3) << 28) |
                                       ((unsigned long long)(L1::MMC + 2) << 32) | ((unsigned long long)(L1::MMC + 4) << 36);
        const unsigned long long CODE = (1ull << 20) | (2ull << 24) | (8ull << 28) | (4ull << 32) | (15ull << 36);
        O.skip(g0 * 2);
        uint4 m = make_uint4(0, 0, 0, 0);
        uint32_t rc4 = 0;
        if (part != 0) {
            const uint32_t rw = refn[q >> 1];
            rc4 = (q & 1) ? (rw & 0xffffu) : (rw >> 16);
            m = ld4(cnt, L1::MATCH, q);
        }
#pragma unroll
        for (int g = g0; g < g1; ++g) {
            const int arr = (int)((ARR >> (4 * g)) & 15u);
            const uint32_t code = (uint32_t)((CODE >> (4 * g)) & 15u);
            uint4 a = ld4(cnt, arr, q);
            if (g == 0) all = lo16(a);
            if (code) a = base_count(a, m, rc4, code);
            O.put(lo16(a));
            O.put(hi16(a));
        }
    } else {
        // groups: 0 reads, 1 matches, 2 mismatches, 3 deletions, 4 insertions, 5 A, 6 C, 7 T, 8 G, 9 N
        // part 0 -> groups 0-2, part 1 -> 3-5, part 2 -> 6-7, part 3 -> 8-9 (contiguous fields)
        const bool S = (STAT == PSS_VARIATION_STRAND);
        const int per = S ? 6 : 2;
        if (TFIN_LOOP && S) {
        // a rolled loop over the groups: the unrolled form is ~600 straight-line instructions per output mode, and
        // thirty warps streaming through it miss the instruction caches (ncu: stall_no_instruction 2.2 per issue)
        const int g0 = part == 0 ? 0 : part == 1 ? 3 : part == 2 ? 6 : 8, g1 = part == 0 ? 3 : part == 1 ? 6 : part == 2 ? 8 : 10;
        // counter array (5 bits) and reference code whose positions also get the matches (4 bits) per group;
        // mismatch classes A, C, G, T, N = 0..4 with 4-bit codes 1, 2, 4, 8, 15; output order A C T G N
        const unsigned long long ARR = (unsigned long long)W_ALL | ((unsigned long long)W_MATCH << 5) | ((unsigned long long)W_MMT << 10) |
                                       ((unsigned long long)W_DEL << 15) | ((unsigned long long)W_INS << 20) | ((unsigned long long)(W_MMC + 0) << 25) |
                                       ((unsigned long long)(W_MMC + 2) << 30) | ((unsigned long long)(W_MMC + 6) << 35) |
                                       ((unsigned long long)(W_MMC + 4) << 40) | ((unsigned long long)(W_MMC + 8) << 45);
        const unsigned long long CODE = (1ull << 20) | (2ull << 24) | (8ull << 28) | (4ull << 32) | (15ull << 36);
        O.skip(g0 * per);
        uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
        uint32_t rc4 = 0;
        if (part != 0) {
            const uint32_t rw = refn[q >> 1];
            rc4 = (q & 1) ? (rw & 0xffffu) : (rw >> 16);
            m0 = ld4(cnt, W_MATCH, q); m1 = ld4(cnt, W_MATCH + 1, q);
        }
#pragma unroll 1
        for (int g = g0; g < g1; ++g) {
            const int arr = (int)((ARR >> (5 * g)) & 31u);
            const uint32_t code = (uint32_t)((CODE >> (4 * g)) & 15u);
            uint4 a0 = ld4(cnt, arr, q), a1 = ld4(cnt, arr + 1, q);
            if (g == 0) { const uint4 sa = a0 + a1; all = lo16(sa) + hi16(sa); }
            if (code) { a0 = base_count(a0, m0, rc4, code); a1 = base_count(a1, m1, rc4, code); }
            O.template put_group<S>(a0, a1);
        }
        return all;
        }
        if (part == 0) {
            const uint4 w0 = ld4(cnt, W_ALL, q), w1 = ld4(cnt, W_ALL + 1, q);
            const uint4 s = w0 + w1;
            all = lo16(s) + hi16(s);
            O.template put_group<S>(w0, w1);
            O.template put_group<S>(ld4(cnt, W_MATCH, q), ld4(cnt, W_MATCH + 1, q));
            O.template put_group<S>(ld4(cnt, W_MMT, q), ld4(cnt, W_MMT + 1, q));
            return all;
        }
        const uint32_t rw = refn[q >> 1];
        const uint32_t rc4 = (q & 1) ? (rw & 0xffffu) : (rw >> 16);
        if (part == 1) {
            O.skip(3 * per);
            O.template put_group<S>(ld4(cnt, W_DEL, q), ld4(cnt, W_DEL + 1, q));
            O.template put_group<S>(ld4(cnt, W_INS, q), ld4(cnt, W_INS + 1, q));
        } else O.skip((part == 2 ? 6 : 8) * per);
        const uint4 m0 = ld4(cnt, W_MATCH, q), m1 = ld4(cnt, W_MATCH + 1, q);
        // mismatch classes A, C, G, T, N = 0..4 with 4-bit codes 1, 2, 4, 8, 15; output order A C T G N
        if (part == 1) {
            O.template put_group<S>(base_count(ld4(cnt, W_MMC + 0, q), m0, rc4, 1u), base_count(ld4(cnt, W_MMC + 1, q), m1, rc4, 1u));
        } else if (part == 2) {
            O.template put_group<S>(base_count(ld4(cnt, W_MMC + 2, q), m0, rc4, 2u), base_count(ld4(cnt, W_MMC + 3, q), m1, rc4, 2u));
            O.template put_group<S>(base_count(ld4(cnt, W_MMC + 6, q), m0, rc4, 8u), base_count(ld4(cnt, W_MMC + 7, q), m1, rc4, 8u));
        } else {
            O.template put_group<S>(base_count(ld4(cnt, W_MMC + 4, q), m0, rc4, 4u), base_count(ld4(cnt, W_MMC + 5, q), m1, rc4, 4u));
            O.template put_group<S>(base_count(ld4(cnt, W_MMC + 8, q), m0, rc4, 15u), base_count(ld4(cnt, W_MMC + 9, q), m1, rc4, 15u));
        }
    }
    return all;
}

template <int STAT, int MODE, int NT>
__device__ __forceinline__ void tile_output(const TileArgs &A, const uint32_t *cnt, const uint32_t *raw,
                                            const uint32_t *refn, int64_t t0l, int tid) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(A.out) & 15) == 0) && ((A.stride & 3) == 0) &&
                         (((t0l - A.start) & 3) == 0);
    // tasks = (sub-part, quad): 4 NQUAD of them (18, 18, 12, 12 fields per quad for variation_strand),
    // dealt to the threads round-robin; a warp always holds one sub-part (NQUAD is a multiple of 32)
    uint32_t seen = 0;
#pragma unroll 1
    for (int task = tid; task < FinParts<STAT>::N * NQUAD; task += NT) {
        const int part = task / NQUAD, q = task - part * NQUAD;
        const int64_t pl = t0l + 4 * q;
        const int64_t rem = A.end - pl;
        if (rem <= 0) continue;
        QuadOut<MODE> O;
        O.p = (MODE == OUT_NARROW8) ? reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(A.out) + (pl - A.start))
              : (MODE == OUT_NARROW) ? reinterpret_cast<int32_t *>(reinterpret_cast<uint16_t *>(A.out) + (pl - A.start))
                                     : A.out + (pl - A.start);
        O.stride = A.stride;
        O.seen = 0;
        O.nvalid = rem > 4 ? 4 : (int)rem;
        // full quads always; partial quads only when the row is padded (stride beyond the range)
        O.vec = aligned && (O.nvalid == 4 || (pl - A.start) + 4 <= A.stride);
        uint4 all = tile_finalize<STAT, MODE>(cnt, refn, q, part, O);
        if (part == 0) {
            if (A.need_raw) all = ld4(raw, 0, q);
            uint8_t *m = A.rowmask + (pl - A.start);
            const bool acc = MODE == OUT_ACCUM;
            if (!acc && O.nvalid == 4 && (reinterpret_cast<uintptr_t>(m) & 3) == 0) {
                *reinterpret_cast<uint32_t *>(m) = (all.x ? 1u : 0u) | (all.y ? 0x100u : 0u) | (all.z ? 0x10000u : 0u) | (all.w ? 0x1000000u : 0u);
            } else {
                if (O.nvalid > 0) m[0] = (all.x ? 1 : 0) | (acc ? m[0] : 0);
                if (O.nvalid > 1) m[1] = (all.y ? 1 : 0) | (acc ? m[1] : 0);
                if (O.nvalid > 2) m[2] = (all.z ? 1 : 0) | (acc ? m[2] : 0);
                if (O.nvalid > 3) m[3] = (all.w ? 1 : 0) | (acc ? m[3] : 0);
            }
        }
        if (MODE == OUT_NARROW8) seen |= O.seen;
    }
    if (MODE == OUT_NARROW8 && (seen >> 8)) atomicOr(A.ovf8, 1);
}

// ------------------------------------------------------------------------------ the kernel

template <int KIND, int STAT, bool MINQ>
__global__ void __launch_bounds__(KindLaunch<KIND>::NT, StatCtas<KIND, STAT>::V)
k_tile(TileArgs A) {
    constexpr int NT = KindLaunch<KIND>::NT, NWARPS = NT / 32;
    constexpr bool IS_VAR = KIND == KIND_VAR || KIND == KIND_VAR1;
    typedef typename KindLay<KIND>::type LV;               // counter layout of the variation kinds
    static_assert(NT % 32 == 0 && NQUAD % 32 == 0, "whole warps per finalisation sub-part");
    constexpr int CH = IS_VAR ? TCH : TCH_COV;             // reads staged per chunk
    static_assert(CH <= NT, "one thread per staged read at least");
    extern __shared__ __align__(128) uint8_t smem[];
    const int NARR = KindInfo<KIND>::NDIFF + KindInfo<KIND>::NPOINT;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);
    uint32_t *raw = cnt + NARR * TT;                                   // exists only when need_raw (tile_smem_bytes)
    RInfo *s_ri = reinterpret_cast<RInfo *>(raw + (A.need_raw ? TT : 0));
    // the rest exists for KIND_VAR only (see tile_smem_bytes)
    uint32_t *s_refn = reinterpret_cast<uint32_t *>(s_ri + CH);        // REFN_WORDS + 4 words
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_refn + REFN_WORDS + 4);   // TSEQ_BYTES + 32
    uint32_t *s_cig = reinterpret_cast<uint32_t *>(s_seq + TSEQ_BYTES + 32);  // TCIG_WORDS + 8
    SpanRec *s_span = reinterpret_cast<SpanRec *>(s_cig + TCIG_WORDS + 8);     // SPAN_MAX records (balanced walk)
    uint16_t *s_ev = reinterpret_cast<uint16_t *>(s_span + (TBAL ? SPAN_MAX : 0));   // EV_SLOTS x EV_STRIDE (DEFER layouts)
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_n;
    __shared__ long long s_seqlo;      // global byte index staged at s_seq[0]
    __shared__ long long s_ciglo;      // global CIGAR word index staged at s_cig[0], -1: not staged
    __shared__ int s_nspan;            // span records of the current chunk

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t t0l = A.start + (int64_t)tile * TT;
#ifdef TILE_TIMING
    unsigned long long ts[8];
#define STAMP(k) do { if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ts[k])); } while (0)
    STAMP(0);
#else
#define STAMP(k)
#endif

    // ---- thread 0 starts the bulk copies of the first chunk before anything else -------------
    TileDesc d;
    uint4 d1 = make_uint4(0, 0, 0, 0), d2 = d1;
    {
        // the whole descriptor in one round trip (thread 0 needs all of it before it can issue the copies)
        const uint4 *dp = reinterpret_cast<const uint4 *>(A.desc + tile);
        const uint4 d0 = __ldg(dp);
        if (IS_VAR && tid == 0) { d1 = __ldg(dp + 1); d2 = __ldg(dp + 2); }
        d.x = (int)d0.x; d.y = (int)d0.y; d.n0 = (int)d0.z;
        d.flags = IS_VAR ? __ldg(A.tflags + tile) : 0u;
    }
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        s_nspan = 0;
        uint32_t bytes = 0, sbytes = 0, cbytes = 0;
        const uint8_t *src = nullptr, *csrc = nullptr;
        const int n = IS_VAR ? d.n0 : min(CH, d.y - d.x);      // (the descriptor's first chunk is sized for the variation kinds)
        if (IS_VAR && n > 0) {
            long long lo = (long long)(((uint64_t)d1.y << 32) | d1.x), hi = (long long)(((uint64_t)d1.w << 32) | d1.z);
            stage_span(A.seq4, lo, hi, 1, &src, &sbytes, &s_seqlo);
            long long clo = (long long)(((uint64_t)d2.y << 32) | d2.x), chi = (long long)(((uint64_t)d2.w << 32) | d2.z);
            s_ciglo = -1;
            if (chi > clo) stage_span(A.cigar, clo, chi, 4, &csrc, &cbytes, &s_ciglo);
        }
        s_n = n;
        if (n > 0) {
            bytes = (uint32_t)n * (uint32_t)sizeof(RInfo) + sbytes + cbytes + (IS_VAR ? (REFN_WORDS + 4) * 4 : 0);
            mbar_arrive_expect_tx(&s_bar, bytes);
            bulk_g2s(s_ri, A.ri + d.x, (uint32_t)n * (uint32_t)sizeof(RInfo), &s_bar);
            if (sbytes) bulk_g2s(s_seq, src, sbytes, &s_bar);
            if (cbytes) bulk_g2s(s_cig, csrc, cbytes, &s_bar);
            if (IS_VAR) bulk_g2s(s_refn, A.refpack + (size_t)tile * REFN_WORDS, (REFN_WORDS + 4) * 4, &s_bar);
        } else mbar_arrive_expect_tx(&s_bar, 0);
    }
    STAMP(1);
#if TPREFETCH > 0
    // CTAs start in tile order, so tile + TPREFETCH starts about when this one ends: the first lane of the last warp
    // (no part in the walk of a typical tile) asks for that tile's records / sequences / CIGARs to be brought into
    // L2, so its bulk copies do not wait for DRAM
    if (tid == NT - 32 && tile + TPREFETCH < (int)gridDim.x) {
        const uint4 *dp = reinterpret_cast<const uint4 *>(A.desc + tile + TPREFETCH);
        const uint4 p0 = __ldg(dp), p1 = __ldg(dp + 1), p2 = __ldg(dp + 2);
        const int pn = (int)p0.z;
        if (pn > 0) {
            const uint8_t *src;
            uint32_t nb;
            long long first;
            bulk_prefetch_l2(A.ri + (int)p0.x, (uint32_t)pn * (uint32_t)sizeof(RInfo));
            if (IS_VAR) {
            const long long lo = (long long)(((uint64_t)p1.y << 32) | p1.x), hi = (long long)(((uint64_t)p1.w << 32) | p1.z);
            stage_span(A.seq4, lo, hi, 1, &src, &nb, &first);
            if (nb) bulk_prefetch_l2(src, nb);
            const long long clo = (long long)(((uint64_t)p2.y << 32) | p2.x), chi = (long long)(((uint64_t)p2.w << 32) | p2.z);
            if (chi > clo) { stage_span(A.cigar, clo, chi, 4, &src, &nb, &first); bulk_prefetch_l2(src, nb); }
            }
        }
    }
#endif

    TileCtx T;
    T.cnt = cnt; T.seqw = reinterpret_cast<const uint32_t *>(s_seq); T.refn = s_refn;
    T.ref = A.ref; T.qual = nullptr; T.ev = s_ev + tid; T.ref_start = A.ref_start; T.ref_len = A.ref_len;
    T.t0 = (int)t0l; T.t1 = T.t0 + TT; T.no_del = A.no_del; T.slow = (d.flags & 1u) != 0;
    uint32_t *rawp = A.need_raw ? raw : nullptr;

    uint32_t parity = 0;
    int cb = d.x;
    bool more = true, staged = false, accum = false;
    int n = 0;
    for (;;) {      // passes of at most PASS_READS reads (one pass unless the tile is extremely deep)
        {   // zero the counters
            const int nw4 = (NARR + (A.need_raw ? 1 : 0)) * TT / 4;
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int k = tid; k < nw4; k += NT) reinterpret_cast<uint4 *>(cnt)[k] = z;
        }
        __syncthreads();           // barrier initialised, counters zeroed
        STAMP(2);
        int in_pass = 0;
        while (more) {
            if (!staged) {
                mbar_wait(&s_bar, parity);
                STAMP(3);
                parity ^= 1;
                n = s_n;
                staged = true;
                if (n == 0) { more = false; break; }
            }
            if (in_pass + n > PASS_READS) break;       // finalise what we have, then come back to this chunk
            if (KIND == KIND_VAR && TBAL && !T.slow) {
                // balanced walk: (A) one thread per read -- interval adds, insertion points, span records
                if (tid < n) {
                    const RInfo r = s_ri[tid];
                    const long long ciglo = s_ciglo;
                    const uint32_t *cg = ciglo >= 0 ? s_cig + (r.cigar_off - ciglo) : A.cigar + r.cigar_off;
                    const uint32_t sn0 = (uint32_t)((r.seq_off - s_seqlo) << 1);
                    if (!(r.cls & CLS_HASSEQ)) {       // every base a mismatch of no class (or, under min_baseq, dropped)
                        if (MINQ) noseq_minq_read<KIND>(T, A, r, rawp);
                        else var_read<LayVar>(T, cg, r, sn0, rawp, 0, 1);
                    } else {
                        if (MINQ) lowq_prefetch(A, r, cb + tid);
#ifndef PSS_EXP_SKIP_WALK
                        var_emit_spans(T, cg, r, sn0, rawp, s_span, &s_nspan);
#endif
                        if (MINQ) var_lowq_fix<LayVar>(T, A, cg, r, sn0, cb + tid, 0, 1);
                    }
                }
                __syncthreads();
                // (B) every warp computes the prefix sums of the spans' word counts itself (cheaper than one more
                // block barrier), then every thread walks the same number of consecutive words
                const int ns = min(s_nspan, SPAN_MAX);
                int total = 0;
                for (int base = 0; base < ns; base += 32) {
                    const int i = base + lane;
                    const int nw = i < ns ? (int)((s_span[i].a_len >> 16) + 7) >> 3 : 0;
                    int inc = nw;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(FULL, inc, o);
                        if (lane >= o) inc += t;
                    }
                    if (i < ns) s_span[i].wbase = (uint32_t)(total + inc - nw);
                    total += __shfl_sync(FULL, inc, 31);
                }
                __syncwarp();
                const int quota = (total + NT - 1) / NT;
                var_walk_words(T, s_span, ns, tid * quota, min(total, (tid + 1) * quota));
            } else if (IS_VAR) {
                // up to four threads share a read: threads [k n, (k + 1) n) are part k of reads 0 .. n-1
                const int nparts = min(4, NT / n);
                const int part = tid / n;
                if (part < nparts) {
                    const RInfo r = s_ri[tid - part * n];
                    const long long ciglo = s_ciglo;
                    const uint32_t *cg = ciglo >= 0 ? s_cig + (r.cigar_off - ciglo) : A.cigar + r.cigar_off;
                    if (MINQ && !(r.cls & CLS_HASSEQ)) {     // quality 0 everywhere: the bases are all dropped
                        if (part == 0) noseq_minq_read<KIND>(T, A, r, rawp);
                    } else {
                        if (MINQ) lowq_prefetch(A, r, cb + tid - part * n);
#ifndef PSS_EXP_SKIP_WALK      // experiment builds only (scripts/build_variant.sh): time of the phases in isolation
                        var_read<LV>(T, cg, r, (uint32_t)((r.seq_off - s_seqlo) << 1), rawp, part, nparts);
#endif
                        if (MINQ) var_lowq_fix<LV>(T, A, cg, r, (uint32_t)((r.seq_off - s_seqlo) << 1), cb + tid - part * n, part, nparts);
                    }
                }
            } else if (tid < n) {
                const RInfo r = s_ri[tid];
                if (MINQ && !(r.cls & CLS_HASSEQ)) noseq_minq_read<KIND>(T, A, r, rawp);
                else {
                    if (MINQ) lowq_prefetch(A, r, cb + tid);
                    cover_read<KIND>(T, A, r, rawp);
                    if (MINQ) cover_lowq_fix<KIND>(T, A, r, cb + tid);
                }
            }
            in_pass += n;
            staged = false;
            cb += n;
            if (cb >= d.y) { more = false; break; }
            __syncthreads();                       // staging buffers free again
            if (tid == 0) {
                int m = min(CH, d.y - cb);
                uint32_t sbytes = 0, cbytes = 0;
                const uint8_t *src = nullptr, *csrc = nullptr;
                if (IS_VAR) {
                    long long lo = A.seq_off[cb], hi = A.seq_off[cb + m];
                    while (m > 1 && hi - lo + 32 > TSEQ_BYTES) { m >>= 1; hi = A.seq_off[cb + m]; }
                    if (hi - lo + 32 > TSEQ_BYTES) { atomicOr(&A.st->toolong, 1); m = 0; }
                    else {
                        stage_span(A.seq4, lo, hi, 1, &src, &sbytes, &s_seqlo);
                        long long clo = A.ri[cb].cigar_off, chi = A.ri[cb + m - 1].cigar_off + A.ri[cb + m - 1].n_cigar;
                        s_ciglo = -1;
                        if (chi > clo && chi - clo + 8 <= TCIG_WORDS) stage_span(A.cigar, clo, chi, 4, &csrc, &cbytes, &s_ciglo);
                    }
                }
                s_n = m;
                s_nspan = 0;
                fence_proxy_async();
                if (m > 0) {
                    mbar_arrive_expect_tx(&s_bar, (uint32_t)m * (uint32_t)sizeof(RInfo) + sbytes + cbytes);
                    bulk_g2s(s_ri, A.ri + cb, (uint32_t)m * (uint32_t)sizeof(RInfo), &s_bar);
                    if (sbytes) bulk_g2s(s_seq, src, sbytes, &s_bar);
                    if (cbytes) bulk_g2s(s_cig, csrc, cbytes, &s_bar);
                } else mbar_arrive_expect_tx(&s_bar, 0);
            }
        }
        STAMP(4);
        __syncthreads();
        STAMP(5);
        {   // prefix sums of the difference arrays (+ raw)
            const int nscan = KindInfo<KIND>::NDIFF + (A.need_raw ? 1 : 0);
            for (int a = warp; a < nscan; a += NWARPS) {
                uint32_t *arr = (a < KindInfo<KIND>::NDIFF) ? cnt + a * TT : raw;
                warp_scan_array(arr, lane);
            }
        }
        __syncthreads();
        STAMP(6);
#ifdef PSS_EXP_SKIP_OUT
        if (A.narrow == 77)
#endif
        if (A.narrow) {
            // 16-bit transport cannot add passes: flag it, the host repeats the call with int32 output
            if (accum || more) { if (tid == 0) atomicOr(&A.st->overflow16, 1); }
            else if (A.narrow == 2) tile_output<STAT, OUT_NARROW8, NT>(A, cnt, raw, s_refn, t0l, tid);
            else tile_output<STAT, OUT_NARROW, NT>(A, cnt, raw, s_refn, t0l, tid);
        } else if (accum) tile_output<STAT, OUT_ACCUM, NT>(A, cnt, raw, s_refn, t0l, tid);
        else tile_output<STAT, OUT_STORE, NT>(A, cnt, raw, s_refn, t0l, tid);
        STAMP(7);
#ifdef TILE_TIMING
        if (tid == 0 && A.timing && (tile & 63) == 0) {
            unsigned long long *o = A.timing + (size_t)(tile >> 6) * 8;
            for (int k = 0; k < 8; ++k) o[k] = ts[k];
        }
#endif
        if (!more) break;
        accum = true;
        __syncthreads();           // everyone is done reading the counters before they are zeroed
    }
}

// ------------------------------------------------------------------ the quality kinds
//
// mapq(_strand), baseq(_strand), baseq_ext(_strand) (opt.pyx:912-1462) when min_baseq == 0.
// Their sums are not all interval adds: max(mapq) and the per-base sum of squared base
// qualities depend on every (read, position) event.  k_tileq splits the work of a tile in two
// kinds of passes over the same staged reads:
//   scatter  one read per thread (baseq_ext: the XOR read walk of the variation kernel):
//            everything constant along a span as +1 / -1 in 32-bit difference arrays per bank
//            (reads, non-deletion events, mapq == 0, mapq^2 as two byte planes), mismatching
//            bases as point updates (count and squared quality);
//   gather   position-owned, no atomics in the loop: a warp owns 128 consecutive positions
//            (4 per lane) and one parity of the reads, walks the reads that reach its window
//            in file order and accumulates max(mapq) / sum(q^2) per bank in registers; the
//            registers are folded into shared-memory arrays once per tile.
// Counters are full 32-bit words (64-bit for the squared sums), so a tile is always a single
// pass whatever its depth: root mean squares are not additive over passes.
// rec() is fused as in k_tile; the fp64 rootmean (opt.pyx:2159-2163) is evaluated per field.

enum { KIND_BQ = 4, KIND_BQX = 5 };          // (3 was mapq on this kernel until k_tilem below)
#ifndef QCH
#define QCH 160                        // reads staged per chunk (a 512-position tile of 30x / 150 bp has ~135 candidates)
#endif
#ifndef BQ_CTAS
#define BQ_CTAS 3
#endif
#define QSEQ_BYTES (QCH * 76 + 64)
#define QQUAL_BYTES (QCH * 152 + 64)
#ifndef QCIG_WORDS
#define QCIG_WORDS 256
#endif

// word arrays of TT (bank = rev | pp << 1):
//   KIND_BQ    n[0..3] nnd[4..7] (difference arrays)  sq64[8..15]
//              baseq (no strand split in its output): TWO banks, pp or not -- n[0..1] nnd[2..3] sq64[4..7] -- half the
//              arrays to zero, scan and finalise, half the gather's accumulators, and 50 KB per CTA: a fourth CTA per SM
//   KIND_BQX   n[0..3] match[4..7] (difference arrays)  mmt[8..11]  sq64[12..19]  sqmm64[20..27]   (LayBqx)
template <int KIND> struct QKind;
// NT: threads per CTA (the first 8 warps gather; baseq_ext wants two threads per staged read in the scatter)
template <> struct QKind<KIND_BQ> { static const int NDIFF = 8, NWORDS = 16, SQ = 8, NT = 256, CTAS = BQ_CTAS; static const bool QUAL = true, SEQ = false, CIG = true; };
template <> struct QKind<KIND_BQX> { static const int NDIFF = 8, NWORDS = 28, SQ = 12, NT = 320, CTAS = 2; static const bool QUAL = true, SEQ = true, CIG = true; };

struct QSmem {       // byte offsets into the dynamic shared memory
    int raw, ri, qoff, refn, seq, qual, cig, total;
};
// banks of a (kind, statistic), CTAs per SM it is compiled for
template <int KIND, int STAT> struct QBanks { static const int NBK = 4, CTAS = QKind<KIND>::CTAS; };
template <> struct QBanks<KIND_BQ, PSS_BASEQ> { static const int NBK = 2, CTAS = 4; };
template <> struct QBanks<KIND_BQX, PSS_BASEQ_EXT> { static const int NBK = 2, CTAS = 3; };     // 14 word arrays: 73 KB per CTA
template <int NBK> struct BqxLay { typedef LayBqx type; };
template <> struct BqxLay<2> { typedef LayBqx2 type; };
template <int KIND, int NBK = 4>
__host__ __device__ constexpr QSmem q_layout() {
    QSmem L{};
    int o = (KIND == KIND_BQ ? 4 * NBK : 7 * NBK) * TT * 4;
    L.raw = o; o += TT * 4;
    L.ri = o; o += QCH * 32;
    L.qoff = o; if (QKind<KIND>::QUAL) o += (QCH + 4) * 8;
    L.refn = o; if (QKind<KIND>::SEQ) o += (REFN_WORDS + 4) * 4;
    L.seq = o; if (QKind<KIND>::SEQ) o += QSEQ_BYTES + 32;
    L.qual = o; if (QKind<KIND>::QUAL) o += QQUAL_BYTES + 32;
    L.cig = o; if (QKind<KIND>::CIG) o += (QCIG_WORDS + 8) * 4;
    L.total = (o + 15) & ~15;
    return L;
}

__device__ __forceinline__ void six32(const uint32_t *b, uint32_t *o6) {   // CountPpStrand order from the banks
    o6[0] = b[0] + b[1] + b[2] + b[3]; o6[1] = b[0] + b[2]; o6[2] = b[1] + b[3]; o6[3] = b[2] + b[3]; o6[4] = b[2]; o6[5] = b[3];
}
__device__ __forceinline__ void six64(const unsigned long long *b, unsigned long long *o6) {
    o6[0] = b[0] + b[1] + b[2] + b[3]; o6[1] = b[0] + b[2]; o6[2] = b[1] + b[3]; o6[3] = b[2] + b[3]; o6[4] = b[2]; o6[5] = b[3];
}
// rootmean (opt.pyx:2159-2163): int(numpy.around(sqrt(sq / n))) in doubles.  Evaluated in integers:
// r is the answer iff (2r - 1)^2 n <= 4 sq <= (2r + 1)^2 n, a tie (sqrt exactly r +- 0.5, which is
// representable, so the reference sees it too) going to the even neighbour.  When the two sides
// of a comparison differ they differ by >= 1 in 4 sq < 2^46, i.e. by more than 2^-46 relative,
// far outside what the two correctly rounded double operations (2^-53 each) can move, so the
// integer decision IS the double-precision result (checked against the double formula on 1.5e8
// cases incl. every tie neighbourhood: scripts/rootmean_check.cpp).  The float guess is within 1.
__device__ __forceinline__ int q_rootmean(unsigned long long sq, uint32_t n) {
    if (n == 0 || sq == 0) return 0;
    const float x = __fdividef((float)sq, (float)n);
    // the mean of squared bytes is < 2^16; anything else (and sums beyond 2^44) takes the double path
    if ((sq >> 44) || x > 1.0e9f) return (int)rint(sqrt((double)sq / (double)n));
    const uint32_t r = (uint32_t)(x * rsqrtf(x) + 0.5f);       // < 2^15: the squares below fit 32 bits
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

// The same value through a float screening: the estimate e of sqrt(sq / n) is off by less than 5e-7 e (conversions, one
// approximate division and one reciprocal square root at 2 ulp each), so |e - rint(e)| <= 0.4999 - 1e-6 e leaves the real
// number less than 0.5 away from rint(e) -- its rounding, with no tie near.  Only estimates next to a half (and sums
// beyond 32 bits) take the integer decision above, out of line: the finalisation of baseq_ext_strand evaluates 18 root
// mean squares per position, and inlining the full form 18 times was a quarter of its instructions.
// (scripts/tlen_fast_check.cpp checks the screening -- its rms column -- against the exact form.)
__device__ __noinline__ int q_rootmean_exact(unsigned long long sq, uint32_t n) { return q_rootmean(sq, n); }
__device__ __forceinline__ int q_rootmean_scr(unsigned long long sq, uint32_t n) {
    const float x = __fdividef((float)(uint32_t)sq, (float)n), e = x * rsqrtf(x), r = rintf(e);   // (sq == 0: NaN, not accepted)
    if (sq == 0) return 0;                                                                          // (n == 0 has sq == 0)
    return ((sq >> 32) == 0 && fabsf(e - r) <= 0.4999f - 1e-6f * e) ? (int)r : q_rootmean_exact(sq, n);
}

// one aligned span [x, x + l) of a read whose first base is query index y: the lanes' four positions
// (minq: events of bases below min_baseq are dropped, opt.pyx:1900-1903; 0 keeps everything)
template <int BANK, bool MINQ, int NBK>
__device__ __forceinline__ void q_gather_span(uint32_t (&acc)[NBK][4], const uint8_t *qual, int p0, int x, int l, int y, uint32_t minq) {
    const int off0 = p0 - x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int off = off0 + i;
        if ((unsigned)off < (unsigned)l) {
            const uint32_t q = qual[y + off];
            if (!MINQ || q >= minq) acc[BANK][i] += q * q;
        }
    }
}

// MINQ (baseq / baseq_strand only): min_baseq > 0 -- a compile-time switch, so the kernels of the default case keep
// their instruction stream (a run-time test in the gather cost baseq 2.93 -> 3.84 ms per 50 Mb)
template <int KIND, int STAT, bool MINQ>
__global__ void __launch_bounds__(QKind<KIND>::NT, QBanks<KIND, STAT>::CTAS)
k_tileq(TileArgs A) {
    using K = QKind<KIND>;
    constexpr int NBK = QBanks<KIND, STAT>::NBK;
    constexpr QSmem L = q_layout<KIND, NBK>();
    // KIND_BQ: n[NBK] nnd[NBK] sq64[NBK]; KIND_BQX: n[NBK] match[NBK] mmt[NBK] sq64[NBK] sqmm64[NBK] (LayBqx / LayBqx2)
    constexpr int K_NDIFF = 2 * NBK, K_NWORDS = KIND == KIND_BQ ? 4 * NBK : 7 * NBK, K_SQ = KIND == KIND_BQ ? 2 * NBK : 3 * NBK;
    typedef typename BqxLay<NBK>::type LX;
    static_assert(KIND == KIND_BQ || (LX::MATCH == NBK && LX::MMT == 2 * NBK && LX::SQMM == 5 * NBK), "LayBqx* offsets");
    constexpr int QNT = K::NT, NWARPS = QNT / 32;
    static_assert(QNT >= 256 && QNT >= QCH, "eight gather warps, one scatter thread per staged read");
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);
    uint32_t *raw = reinterpret_cast<uint32_t *>(smem + L.raw);
    RInfo *s_ri = reinterpret_cast<RInfo *>(smem + L.ri);
    long long *s_qoff = reinterpret_cast<long long *>(smem + L.qoff);
    uint32_t *s_refn = reinterpret_cast<uint32_t *>(smem + L.refn);
    uint8_t *s_seq = smem + L.seq;
    uint8_t *s_qual = smem + L.qual;
    uint32_t *s_cig = reinterpret_cast<uint32_t *>(smem + L.cig);
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_n;
    __shared__ long long s_seqlo, s_quallo, s_qofflo, s_ciglo;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t t0l = A.start + (int64_t)tile * TT;
    int dx, dy;
    uint32_t dflags = 0;
    {
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(A.desc + tile));
        dx = (int)d0.x; dy = (int)d0.y;
        if (K::SEQ) dflags = __ldg(A.tflags + tile);
    }
    // thread 0 sizes a chunk (sequence / quality spans must fit the staging buffers) and issues its bulk copies
    auto issue = [&](int cb, bool first) {
        int n = min(QCH, dy - cb);
        uint32_t b1 = 0, b2 = 0, b3 = 0, b4 = 0;
        const uint8_t *p1 = nullptr, *p2 = nullptr, *p3 = nullptr, *p4 = nullptr;
        long long slo = 0, shi = 0, qlo = 0, qhi = 0;
        while (n > 0) {
            if (K::SEQ) { slo = A.seq_off[cb]; shi = A.seq_off[cb + n]; }
            if (K::QUAL) { qlo = A.qual_off[cb]; qhi = A.qual_off[cb + n]; }
            if (shi - slo + 32 <= QSEQ_BYTES && qhi - qlo + 32 <= QQUAL_BYTES) break;
            if (n == 1) { atomicOr(&A.st->toolong, 1); n = 0; }
            else n >>= 1;
        }
        s_ciglo = -1;
        if (n > 0) {
            if (K::SEQ) stage_span(A.seq4, slo, shi, 1, &p1, &b1, &s_seqlo);
            if (K::QUAL) {
                stage_span(A.qual, qlo, qhi, 1, &p2, &b2, &s_quallo);
                stage_span(A.qual_off, (long long)cb, (long long)cb + n, 8, &p3, &b3, &s_qofflo);
            }
            if (K::CIG) {
                const long long clo = A.ri[cb].cigar_off, chi = A.ri[cb + n - 1].cigar_off + A.ri[cb + n - 1].n_cigar;
                if (chi > clo && chi - clo + 8 <= QCIG_WORDS) stage_span(A.cigar, clo, chi, 4, &p4, &b4, &s_ciglo);
            }
        }
        s_n = n;
        fence_proxy_async();
        if (n > 0) {
            const uint32_t rb = (K::SEQ && first) ? (REFN_WORDS + 4) * 4 : 0;
            mbar_arrive_expect_tx(&s_bar, (uint32_t)n * (uint32_t)sizeof(RInfo) + b1 + b2 + b3 + b4 + rb);
            bulk_g2s(s_ri, A.ri + cb, (uint32_t)n * (uint32_t)sizeof(RInfo), &s_bar);
            if (b1) bulk_g2s(s_seq, p1, b1, &s_bar);
            if (b2) bulk_g2s(s_qual, p2, b2, &s_bar);
            if (b3) bulk_g2s(s_qoff, p3, b3, &s_bar);
            if (b4) bulk_g2s(s_cig, p4, b4, &s_bar);
            if (rb) bulk_g2s(s_refn, A.refpack + (size_t)tile * REFN_WORDS, rb, &s_bar);
        } else mbar_arrive_expect_tx(&s_bar, 0);
    };
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        issue(dx, true);
    }
#if TPREFETCH > 0
    if (tid == QNT - 32 && tile + TPREFETCH < (int)gridDim.x) {      // a later tile's records (+ qualities) into L2, as in k_tile
        const uint4 p0 = __ldg(reinterpret_cast<const uint4 *>(A.desc + tile + TPREFETCH));
        const int px = (int)p0.x, pn = min(QCH, (int)p0.y - px);
        if (pn > 0) {
            bulk_prefetch_l2(A.ri + px, (uint32_t)pn * (uint32_t)sizeof(RInfo));
            if (K::QUAL) {
                const uint8_t *src;
                uint32_t nb;
                long long first;
                stage_span(A.qual, A.qual_off[px], A.qual_off[px + pn], 1, &src, &nb, &first);
                if (nb) bulk_prefetch_l2(src, nb);
            }
        }
    }
#endif
    {   // zero the counters (+ raw)
        const int nw4 = (K_NWORDS + 1) * TT / 4;
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < nw4; k += QNT) reinterpret_cast<uint4 *>(cnt)[k] = z;
    }
    TileCtx T;
    T.cnt = cnt; T.seqw = reinterpret_cast<const uint32_t *>(s_seq); T.refn = s_refn;
    T.ref = A.ref; T.qual = s_qual; T.ev = nullptr; T.ref_start = A.ref_start; T.ref_len = A.ref_len;
    T.t0 = (int)t0l; T.t1 = T.t0 + TT; T.no_del = A.no_del; T.slow = (dflags & 1u) != 0;
    uint32_t *rawp = A.need_raw ? raw : nullptr;
    __syncthreads();           // barrier initialised, counters zeroed
    const int maxspan = A.st->maxspan;
    constexpr bool minq = MINQ;                            // min_baseq > 0: the marked bases' events are dropped
    const uint32_t gminq = MINQ ? (uint32_t)A.min_baseq : 0u;

    // gather state: this warp's window and read parity; per bank, per owned position
    const int gw0 = T.t0 + (warp & 3) * 128, gw1 = gw0 + 128, gpar = warp >> 2;
    const int gp0 = gw0 + 4 * lane;
    uint32_t acc[NBK][4];
#pragma unroll
    for (int b = 0; b < NBK; ++b)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[b][i] = 0;

    uint32_t parity = 0;
    for (int cb = dx; cb < dy;) {
        mbar_wait(&s_bar, parity);
        parity ^= 1;
        const int n = s_n;
        if (n == 0) break;
        const long long ciglo = s_ciglo, quallo = K::QUAL ? s_quallo : 0;
        const int qskew = K::QUAL ? (int)(cb - s_qofflo) : 0;
        // ---- scatter ------------------------------------------------------------------------
        if (KIND == KIND_BQX) {
            const int nparts = min(4, QNT / n);
            const int part = tid / n;
            if (part < nparts) {
                const int j = tid - part * n;
                const RInfo r = s_ri[j];
                const uint32_t *cg = ciglo >= 0 ? s_cig + (r.cigar_off - ciglo) : A.cigar + r.cigar_off;
                var_read<LX>(T, cg, r, (uint32_t)((r.seq_off - s_seqlo) << 1), rawp, part, nparts,
                                 (int)(s_qoff[j + qskew] - quallo));
            }
        } else if (tid < n) {
            const RInfo r = s_ri[tid];
            const uint32_t cls = r.cls;
            if (cls & CLS_ADMITTED) {
                if (rawp) interval_add(rawp, r.pos, r.end - r.pos, T.t0, 1u);
                if (cls & CLS_COUNTED) {
                    const int bank = NBK == 2 ? (int)((cls >> 1) & 1u) : (int)(cls & 3u);
                    if (minq && !(cls & CLS_HASSEQ)) {
                        // KIND_BQ under min_baseq, no sequence: quality 0 at every base -- only the deletion / ref-skip events remain
                        if (!T.no_del) {
                            if (cls & CLS_NOREF) interval_add(cnt + bank * TT, r.pos, 1, T.t0, 1u);
                            else {
                                const uint32_t *cg = ciglo >= 0 ? s_cig + (r.cigar_off - ciglo) : A.cigar + r.cigar_off;
                                int x = r.pos;
                                for (int k = 0; k < (int)r.n_cigar && x < T.t1; ++k) {
                                    const uint32_t w = cg[k];
                                    const int op = w & 15, l = (int)(w >> 4);
                                    if (op_is_match(op)) x += l;
                                    else if (op == OP_D || op == OP_N) { if (x + l > T.t0) interval_add(cnt + bank * TT, x, l, T.t0, 1u); x += l; }
                                }
                            }
                        }
                    } else {                       // KIND_BQ: n counts every event, nnd the non-deletions (opt.pyx:1108-1112)
                        if (!T.no_del) interval_add(cnt + bank * TT, r.pos, r.end - r.pos, T.t0, 1u);
                        // under min_baseq the marked bases (k_lowq) are taken back out of both counts
                        const uint32_t *mw = minq ? A.lowq + (r.seq_off >> 4) + (cb + tid) : nullptr;
                        auto undo = [&](int p, int run) {
                            if (!T.no_del) interval_add(cnt + bank * TT, T.t0 + p, run, T.t0, 0u - 1u);
                            interval_add(cnt + (NBK + bank) * TT, T.t0 + p, run, T.t0, 0u - 1u);
                        };
                        if (cls & CLS_SIMPLE) {
                            interval_add(cnt + (NBK + bank) * TT, r.pos, r.end - r.pos, T.t0, 1u);
                            if (minq) lowq_runs(mw, r.pos, r.end - r.pos, 0, T.t0, undo);
                        } else if (!(cls & CLS_NOREF)) {
                            const uint32_t *cg = ciglo >= 0 ? s_cig + (r.cigar_off - ciglo) : A.cigar + r.cigar_off;
                            const int nc = (int)r.n_cigar;
                            int x = r.pos, y = 0;
                            for (int k = 0; k < nc && x < T.t1; ++k) {
                                const uint32_t w = cg[k];
                                const int op = w & 15, l = (int)(w >> 4);
                                if (op_is_match(op)) {
                                    if (x + l > T.t0) {
                                        interval_add(cnt + (NBK + bank) * TT, x, l, T.t0, 1u);
                                        if (minq) lowq_runs(mw, x, l, y, T.t0, undo);
                                    }
                                    x += l; y += l;
                                } else if (op == OP_D || op == OP_N) x += l;
                                else if (op == OP_I || op == OP_S) y += l;
                            }
                        }
                    }
                }
            }
        }
        // the lanes of a warp leave the scatter at different times; the gather below is warp-uniform
        // work and must run once per warp, not once per group of lanes that drifted apart
        __syncwarp();
        // ---- gather -------------------------------------------------------------------------
        if (warp < 8) {
            int j;
            {   // first read of the chunk that can reach this warp's window
                int lo = 0, hi = n;
                const int key = gw0 - maxspan + 1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_ri[mid].pos < key) lo = mid + 1; else hi = mid; }
                j = lo + (((lo ^ cb) & 1) != gpar);     // parity of the read's index in the batch
            }
            // the record of the next read is loaded before the current one is processed (its latency
            // hides behind the span work); s_ri has room for the read past the last one.  (Decoding 32 reads once, one
            // per lane, and broadcasting the four words a read needs by shuffles was measured: fewer instructions, but
            // the shuffle -> compare -> branch chain is exposed -- baseq 2.84 -> 2.90 ms, baseq_ext 5.69 -> 6.20 ms.)
            uint4 h = *reinterpret_cast<const uint4 *>(&s_ri[min(j, QCH - 1)]);     // pos, end, cls, n_cigar
            long long qo = K::QUAL ? s_qoff[min(j, QCH - 1) + qskew] : 0;
            for (; j < n; j += 2) {
                const int rpos = (int)h.x, rend = (int)h.y, nc = (int)h.w;
                const uint32_t cls = h.z;
                const long long qoff = qo;
                const int jj = j;
                {
                    const int jn = min(j + 2, QCH - 1);
                    h = *reinterpret_cast<const uint4 *>(&s_ri[jn]);
                    if (K::QUAL) qo = s_qoff[jn + qskew];
                }
                if (rpos >= gw1) break;
                const int bank = NBK == 2 ? (int)((cls >> 1) & 1u) : (int)(cls & 3u);
                {
                    // counted, with sequence (else quality 0) and with bases at all
                    if (rend <= gw0 || (cls & (CLS_COUNTED | CLS_HASSEQ | CLS_NOREF)) != (CLS_COUNTED | CLS_HASSEQ)) continue;
                    const uint8_t *qual = s_qual + (int)(qoff - quallo);
#define QSPAN(B, x, l, y) q_gather_span<B, MINQ>(acc, qual, gp0, x, l, y, gminq)
#define QSPANS(x, l, y) do { if (NBK == 2) { if (bank) QSPAN(NBK - 1, x, l, y); else QSPAN(0, x, l, y); } \
                             else switch (bank) { case 0: QSPAN(0, x, l, y); break; case 1: QSPAN(1, x, l, y); break; \
                                                  case 2: QSPAN(NBK - 2, x, l, y); break; default: QSPAN(NBK - 1, x, l, y); break; } } while (0)
                    if (cls & CLS_SIMPLE) QSPANS(rpos, rend - rpos, 0);
                    else {
                        const long long co = (long long)(((uint64_t)s_ri[jj].cigar_off));
                        const uint32_t *cg = ciglo >= 0 ? s_cig + (co - ciglo) : A.cigar + co;
                        int x = rpos, y = 0;
                        for (int k = 0; k < nc && x < gw1; ++k) {
                            const uint32_t w = cg[k];
                            const int op = w & 15, l = (int)(w >> 4);
                            if (op_is_match(op)) {
                                if (x + l > gw0) QSPANS(x, l, y);
                                x += l; y += l;
                            } else if (op == OP_D || op == OP_N) x += l;
                            else if (op == OP_I || op == OP_S) y += l;
                        }
                    }
#undef QSPANS
#undef QSPAN
                }
            }
        }
        cb += n;
        if (cb >= dy) break;
        __syncthreads();                       // staging buffers free again
        if (tid == 0) issue(cb, false);
    }
    // fold the gather registers into the shared arrays (two warps own each position)
    if (warp < 8) {
        const int pi = (warp & 3) * 128 + 4 * lane;
#pragma unroll
        for (int b = 0; b < NBK; ++b)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (acc[b][i]) {
                    atomicAdd(reinterpret_cast<unsigned long long *>(cnt + K_SQ * TT) + b * TT + pi + i,
                                   (unsigned long long)acc[b][i]);
                }
    }
    __syncthreads();
    {   // prefix sums of the difference arrays (+ raw)
        const int nscan = K_NDIFF + (A.need_raw ? 1 : 0);
        for (int a = warp; a < nscan; a += NWARPS) warp_scan_array(a < K_NDIFF ? cnt + a * TT : raw, lane);
    }
    __syncthreads();
    // ---- rec(): one position per thread and iteration, 128-byte rows per warp and field ------------
    const unsigned long long *sq64 = reinterpret_cast<const unsigned long long *>(cnt + K_SQ * TT);
    const int64_t st = A.stride;
#pragma unroll 1
    for (int p = tid; p < TT; p += QNT) {
        const int64_t pl = t0l + p;
        if (pl >= A.end) break;
        int32_t *o = A.out + (pl - A.start);
        if (NBK == 2 && KIND == KIND_BQ) {
            // baseq: banks not-pp, pp -> reads_all, reads_pp, rms_baseq, rms_baseq_pp (n counts every event, nnd the
            // non-deletions, opt.pyx:1108-1112; under no_del the deletions are not counted at all)
            const uint32_t d0 = cnt[NBK * TT + p], d1 = cnt[(NBK + 1) * TT + p];
            const uint32_t n0 = A.no_del ? d0 : cnt[p], n1 = A.no_del ? d1 : cnt[TT + p];
            const unsigned long long s0 = sq64[p], s1 = sq64[TT + p];
            o[0] = n0 + n1; o[st] = n1; o[2 * st] = q_rootmean_scr(s0 + s1, d0 + d1); o[3 * st] = q_rootmean_scr(s1, d1);
            A.rowmask[pl - A.start] = (A.need_raw ? raw[p] : n0 + n1) ? 1 : 0;
        } else if (NBK == 2) {
            // baseq_ext: banks not-pp, pp -> reads, matches, mismatches, rms_baseq, rms_baseq_matches, rms_baseq_mismatches,
            // each all / pp (opt.pyx:1239-1330)
            const unsigned long long *sqmm = reinterpret_cast<const unsigned long long *>(cnt + LX::SQMM * TT);
            uint32_t nn[2], mt[2], mm[2];
            unsigned long long sq[2], sm[2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                nn[b] = cnt[b * TT + p]; mt[b] = cnt[(LX::MATCH + b) * TT + p]; mm[b] = cnt[(LX::MMT + b) * TT + p];
                sq[b] = sq64[b * TT + p]; sm[b] = sqmm[b * TT + p];
            }
#pragma unroll
            for (int jx = 0; jx < 2; ++jx) {          // all = both banks, pp = bank 1
                const uint32_t N = jx ? nn[1] : nn[0] + nn[1], Mn = jx ? mt[1] : mt[0] + mt[1], Mm = jx ? mm[1] : mm[0] + mm[1];
                const unsigned long long S = jx ? sq[1] : sq[0] + sq[1], Sm = jx ? sm[1] : sm[0] + sm[1];
                o[jx * st] = N; o[(2 + jx) * st] = Mn; o[(4 + jx) * st] = Mm;
                o[(6 + jx) * st] = q_rootmean_scr(S, Mn + Mm); o[(8 + jx) * st] = q_rootmean_scr(S - Sm, Mn);
                o[(10 + jx) * st] = q_rootmean_scr(Sm, Mm);
            }
            A.rowmask[pl - A.start] = (A.need_raw ? raw[p] : nn[0] + nn[1]) ? 1 : 0;
        } else {
        uint32_t nb[4], N[6];
#pragma unroll
        for (int b = 0; b < 4; ++b) nb[b] = cnt[b * TT + p];
        {
            uint32_t db[4], D[6];
            unsigned long long sb[4], S[6];
#pragma unroll
            for (int b = 0; b < 4; ++b) sb[b] = sq64[b * TT + p];
            six64(sb, S);
            if (KIND == KIND_BQ) {
#pragma unroll
                for (int b = 0; b < 4; ++b) db[b] = cnt[(4 + b) * TT + p];
                if (A.no_del) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) nb[b] = db[b];
                }
                six32(nb, N); six32(db, D);
                if (STAT == PSS_BASEQ) {
                    o[0] = N[0]; o[st] = N[3]; o[2 * st] = q_rootmean_scr(S[0], D[0]); o[3 * st] = q_rootmean_scr(S[3], D[3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k) { o[k * st] = N[k]; o[(6 + k) * st] = q_rootmean_scr(S[k], D[k]); }
                }
            } else {
                uint32_t mtb[4], mmb[4], Mn[6], Mm[6];
                unsigned long long smb[4], Sm[6];
                const unsigned long long *sqmm = reinterpret_cast<const unsigned long long *>(cnt + LayBqx::SQMM * TT);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    mtb[b] = cnt[(LayBqx::MATCH + b) * TT + p]; mmb[b] = cnt[(LayBqx::MMT + b) * TT + p];
                    smb[b] = sqmm[b * TT + p];
                }
                six32(nb, N); six32(mtb, Mn); six32(mmb, Mm); six64(smb, Sm);     // Sm: squared qualities of the mismatches
                if (STAT == PSS_BASEQ_EXT) {
                    const int idx[2] = {0, 3};
#pragma unroll
                    for (int jx = 0; jx < 2; ++jx) {
                        const int k = idx[jx];
                        o[jx * st] = N[k]; o[(2 + jx) * st] = Mn[k]; o[(4 + jx) * st] = Mm[k];
                        o[(6 + jx) * st] = q_rootmean_scr(S[k], Mn[k] + Mm[k]); o[(8 + jx) * st] = q_rootmean_scr(S[k] - Sm[k], Mn[k]);
                        o[(10 + jx) * st] = q_rootmean_scr(Sm[k], Mm[k]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        o[k * st] = N[k]; o[(6 + k) * st] = Mn[k]; o[(12 + k) * st] = Mm[k];
                        o[(18 + k) * st] = q_rootmean_scr(S[k], Mn[k] + Mm[k]); o[(24 + k) * st] = q_rootmean_scr(S[k] - Sm[k], Mn[k]);
                        o[(30 + k) * st] = q_rootmean_scr(Sm[k], Mm[k]);
                    }
                }
            }
        }
        A.rowmask[pl - A.start] = (A.need_raw ? raw[p] : N[0]) ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------ mapq, mapq_strand: k_tilem
//
// Mapq / MapqStrand (opt.pyx:912-1100) for ANY min_baseq / no_del.  Everything MapqHelper keeps is constant along the
// intervals in which a read contributes events -- the whole read (M, D and N spans alike), its aligned spans only
// under no_del, the runs of bases at or above the threshold under min_baseq -- so n, nz and sum(mapq^2) are interval
// adds as in the coverage kinds, and max(mapq) is an interval MAXIMUM: at most four shared-memory atomic maxima per
// interval into the block structure of pss_imax.cuh, resolved once per tile.  No per-position gather is left
// (k_tileq<KIND_MAPQ> walked every read once per 128-position window it reaches: 21 thousand warp instructions per
// tile for mapq_strand, about two thirds of them in the gather).
//
// Counters (NBK banks: rev | pp << 1 for mapq_strand, pp for mapq), words of shared memory:
//   a tile with at most 65535 candidate reads ("shallow": every tile of ordinary sequencing depth)
//       [bank][TT]        n | nz << 16   -- neither half can exceed the candidate count
//       [NBK + bank][TT]  sum(mapq^2)    -- <= 65535 * 255^2 < 2^32
//   a deeper tile: the same 2 NBK TT words hold 4 NBK arrays of TT / 2 positions (n, nz, the low byte plane and the high
//       plane of sum(mapq^2), 32-bit each, as k_tileq kept them) and the tile is done in two sweeps over its reads,
//       one per half of its positions -- root mean squares are not additive over passes of reads, but the positions
//       of a tile are independent of each other;
//   sfx / pfx / pt [bank][TT], tab [level][bank][32]: the interval maxima, always over the whole tile (first sweep).
// mapq_strand: banks rev | pp << 1, 256 threads x 4 CTAs per SM (47 KB of shared memory each).  mapq reports all / pp only:
// TWO banks (pp or not) -- half the arrays to zero, scan and resolve, 27 KB -- and 128 threads x 8 CTAs per SM: the phases
// of a tile are short and separated by block barriers, so more, smaller CTAs cover each other's barriers better (ncu of
// the four-bank form on mapq: 2.9 warps per issue waiting at barriers, 44 % of the warp slots active).
template <int STAT> struct MqCfg { static const int NBK = 4, NT = 256, CTAS = 4, CH = 160; };     // CH: reads staged per chunk
template <> struct MqCfg<PSS_MAPQ> { static const int NBK = 2, NT = 128, CTAS = 8, CH = 136; };  // (a tile of 30x / 150 bp has ~135 candidates)
#define MQ_NB (TT / IMAX_BLK)          // blocks per tile: one lane each in the resolution
#define MQ_NLEV 6                      // table levels 0 .. log2(MQ_NB)
#define MQ_PS (TT + TT / 8)            // = imax_padded(TT): words of one bank's sfx / pfx / pt array (position p at imax_idx(p))
static_assert(MQ_NB == 32 && imax_levels(MQ_NB) == MQ_NLEV && MQ_PS == imax_padded(TT), "one warp resolves a bank: 32 blocks of 16 positions");
// words of shared memory for NBK banks: cnt 2 NBK TT | sfx NBK PS | pfx NBK PS | pt NBK PS | tab [level][bank][32]
__host__ __device__ constexpr int mq_words(int nbk) { return 2 * nbk * TT + 3 * nbk * MQ_PS + MQ_NLEV * nbk * MQ_NB; }

// shared-memory reductions by explicit state space: mq_emit below is deliberately NOT inlined (one copy instead of one per
// CIGAR case of the read walk), and through a generic pointer the compiler would emit generic atomics
__device__ __forceinline__ void sred_add(uint32_t *p, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
struct AtomicMaxOp {
    __device__ __forceinline__ void operator()(uint32_t *p, uint32_t v) const {
        asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
    }
};

// One interval of events of a read -- reference positions [x, x + l) -- into the counters of the sweep's window
// [h0, h0 + tw) and (first sweep) into the tile's interval maxima.  flags: deep | first sweep << 1 | bank << 2.
template <int NBK>
__device__ __noinline__ void mq_emit(uint32_t *cnt, int tw, int h0, int t0, uint32_t flags, uint32_t q, int x, int l) {
    const bool deep = flags & 1u;
    const int bank = (int)(flags >> 2);
    const int a = max(x - h0, 0), b = min(x + l - h0, tw);
    if (a < b) {
        uint32_t *p = cnt + bank * tw;
        const uint32_t sq = q * q;
        if (!deep) {
            const uint32_t inc = q ? 1u : 0x10001u;              // n | nz << 16
            sred_add(p + a, inc);
            if (b < tw) sred_add(p + b, 0u - inc);
            if (q) {
                p += NBK * tw;
                sred_add(p + a, sq);
                if (b < tw) sred_add(p + b, 0u - sq);
            }
        } else {
            sred_add(p + a, 1u);
            if (b < tw) sred_add(p + b, 0u - 1u);
            if (!q) {
                p += NBK * tw;
                sred_add(p + a, 1u);
                if (b < tw) sred_add(p + b, 0u - 1u);
            } else {
                const uint32_t lo = sq & 255u, hi = sq >> 8;
                if (lo) { uint32_t *pl = p + 2 * NBK * tw; sred_add(pl + a, lo); if (b < tw) sred_add(pl + b, 0u - lo); }
                if (hi) { uint32_t *ph = p + 3 * NBK * tw; sred_add(ph + a, hi); if (b < tw) sred_add(ph + b, 0u - hi); }
            }
        }
    }
    if ((flags & 2u) && q) {
        const int a2 = max(x - t0, 0), b2 = min(x + l - t0, TT);
        if (a2 < b2) {
            uint32_t *sfx = cnt + 2 * NBK * TT + bank * MQ_PS;
            imax_update(a2, b2, q, sfx, sfx + NBK * MQ_PS, sfx + 2 * NBK * MQ_PS, cnt + 2 * NBK * TT + 3 * NBK * MQ_PS + bank * MQ_NB,
                        NBK * MQ_NB, AtomicMaxOp());
        }
    }
}
__device__ __forceinline__ uint4 max4(uint4 a, uint4 b) { return make_uint4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }

// Four positions of a bank's counters (cq = cnt + 4 quad; counter arrays are tw words apart): n, nz, and the root mean
// square over the banks of MASK (four banks, CountPpStrand: all 0b1111, fwd 0b0101, rev 0b1010, pp 0b1100, pp_fwd 0b0100,
// pp_rev 0b1000; two banks: all 0b11, pp 0b10).
__device__ __forceinline__ uint4 mq_ldn(const uint32_t *cq, int b, int tw, bool deep) {
    const uint4 w = *reinterpret_cast<const uint4 *>(cq + b * tw);
    return deep ? w : lo16(w);
}
template <int NBK>
__device__ __forceinline__ uint4 mq_ldz(const uint32_t *cq, int b, int tw, bool deep) {
    return deep ? *reinterpret_cast<const uint4 *>(cq + (NBK + b) * tw) : hi16(*reinterpret_cast<const uint4 *>(cq + b * tw));
}
// rootmean (opt.pyx:2159-2163) of a shallow tile: n <= 65535 and sq <= 255^2 n < 2^32, so the answer is at most 255.
// A float estimate of sqrt(sq / n) is off by < 2e-4 there (two approximate operations, 2 ulp each, on values <= 255), so
// unless it lies within 1e-3 of a half its nearest integer IS the answer, and only the (rare) estimates next to a half
// take the exact integer decision, out of line.  The first version inlined q_rootmean -- two 64-bit boundary checks and a
// copy of its double-precision fallback -- at each of the 24 places of the finalisation: 5088 SASS instructions, ncu
// showed 3.1 warps per issue waiting for the instruction cache and 37 % of the executed instructions in those checks.
// (scripts/mq_root_check.cpp runs this logic against the reference's double formula around every boundary.)
__device__ __noinline__ uint32_t mq_root_tie(uint32_t sq, uint32_t n, uint32_t r, bool above) {
    // the estimate lies in [r, r + 0.5] or in [r - 0.5, r): only the boundary on that side can be crossed (r >= 1 below).
    // Boundary c / 2 with c = 2 r +- 1:  sqrt(sq / n) <=> c / 2  is  4 sq <=> c^2 n  is  sq <=> t (+ rem / 4); a tie (equality,
    // which the reference's correctly rounded double operations see too) goes to the even neighbour
    const uint32_t c = above ? 2 * r + 1 : 2 * r - 1;
    const unsigned long long cn = (unsigned long long)(c * c) * n;             // c^2 <= 511^2, n < 2^16: cn / 4 < 2^32
    const uint32_t t = (uint32_t)(cn >> 2);
    const bool exact = ((uint32_t)cn & 3u) == 0, odd = r & 1u, eq = sq == t;
    const bool up = above & ((sq > t) | (eq & exact & odd));
    const bool dn = !above & ((sq < t) | (eq & (!exact | odd)));
    return r + (up ? 1u : 0u) - (dn ? 1u : 0u);
}
__device__ __forceinline__ uint32_t mq_root32(uint32_t sq, uint32_t n) {
    const float x = __fdividef((float)sq, (float)n);
    const float rf = x * rsqrtf(x);                         // (sq == 0: NaN, discarded below; n == 0 has sq == 0)
    const float rr = rintf(rf);
    uint32_t r = (uint32_t)rr;
    if (fabsf(rf - rr) > 0.499f) r = mq_root_tie(sq, n, r, rf >= rr);
    return sq ? r : 0u;
}
__device__ __noinline__ uint32_t mq_root_any(unsigned long long sq, uint32_t n) { return (uint32_t)q_rootmean(sq, n); }

template <int MASK, int NBK>
__device__ __forceinline__ uint4 mq_rms(const uint32_t *cq, int tw, bool deep) {
    uint4 n = make_uint4(0, 0, 0, 0);
    if (!deep) {
        // all the reads of the tile together have sum(mapq^2) <= 65535 * 255^2 < 2^32: word sums over the banks
        uint4 s = n;
#pragma unroll
        for (int b = 0; b < NBK; ++b) {
            if (!((MASK >> b) & 1)) continue;
            n = n + lo16(*reinterpret_cast<const uint4 *>(cq + b * tw));
            s = s + *reinterpret_cast<const uint4 *>(cq + (NBK + b) * tw);
        }
        return make_uint4(mq_root32(s.x, n.x), mq_root32(s.y, n.y), mq_root32(s.z, n.z), mq_root32(s.w, n.w));
    }
    unsigned long long s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 1
    for (int b = 0; b < NBK; ++b) {
        if (!((MASK >> b) & 1)) continue;
        n = n + *reinterpret_cast<const uint4 *>(cq + b * tw);
        const uint4 l = *reinterpret_cast<const uint4 *>(cq + (2 * NBK + b) * tw), h = *reinterpret_cast<const uint4 *>(cq + (3 * NBK + b) * tw);
        s0 += l.x + ((unsigned long long)h.x << 8); s1 += l.y + ((unsigned long long)h.y << 8);
        s2 += l.z + ((unsigned long long)h.z << 8); s3 += l.w + ((unsigned long long)h.w << 8);
    }
    return make_uint4(mq_root_any(s0, n.x), mq_root_any(s1, n.y), mq_root_any(s2, n.z), mq_root_any(s3, n.w));
}

template <int STAT, bool MINQ>
__global__ void __launch_bounds__(MqCfg<STAT>::NT, MqCfg<STAT>::CTAS)
k_tilem(TileArgs A) {
    constexpr bool STRAND = STAT == PSS_MAPQ_STRAND;
    constexpr int NBK = MqCfg<STAT>::NBK, MQ_NT = MqCfg<STAT>::NT, MQ_CH = MqCfg<STAT>::CH, MQ_WORDS = mq_words(NBK);
    constexpr int NWARPS = MQ_NT / 32;
    static_assert(NWARPS == 2 * NBK && (NWARPS & (NWARPS - 1)) == 0, "a warp per counter array in the scans, a warp per bank in the resolution");
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);
    uint32_t *sfx = cnt + 2 * NBK * TT, *pfx = sfx + NBK * MQ_PS, *pt = pfx + NBK * MQ_PS, *tab = pt + NBK * MQ_PS;
    uint32_t *raw = cnt + MQ_WORDS;                                          // exists only when need_raw
    RInfo *s_ri = reinterpret_cast<RInfo *>(raw + (A.need_raw ? TT : 0));
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_n;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t t0l = A.start + (int64_t)tile * TT;
    const int t0 = (int)t0l;
    int dx, dy;
    {
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(A.desc + tile));
        dx = (int)d0.x; dy = (int)d0.y;
    }
    const bool deep = dy - dx > 65535;
    const int tw = deep ? TT / 2 : TT;                   // positions per sweep = length of a counter array
    auto issue = [&](int cb) {
        const int n = min(MQ_CH, dy - cb);
        s_n = n;
        fence_proxy_async();
        if (n > 0) {
            mbar_arrive_expect_tx(&s_bar, (uint32_t)n * (uint32_t)sizeof(RInfo));
            bulk_g2s(s_ri, A.ri + cb, (uint32_t)n * (uint32_t)sizeof(RInfo), &s_bar);
        } else mbar_arrive_expect_tx(&s_bar, 0);
    };
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        issue(dx);
    }
#if TPREFETCH > 0
    if (tid == MQ_NT - 32 && tile + TPREFETCH < (int)gridDim.x) {      // a later tile's records into L2, as in k_tile
        const uint4 p0 = __ldg(reinterpret_cast<const uint4 *>(A.desc + tile + TPREFETCH));
        const int px = (int)p0.x, pn = min(MQ_CH, (int)p0.y - px);
        if (pn > 0) bulk_prefetch_l2(A.ri + px, (uint32_t)pn * (uint32_t)sizeof(RInfo));
    }
#endif
    {   // zero the counters, the maxima (+ raw)
        const int nw4 = (MQ_WORDS + (A.need_raw ? TT : 0)) / 4;
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < nw4; k += MQ_NT) reinterpret_cast<uint4 *>(cnt)[k] = z;
    }
    uint32_t *rawp = A.need_raw ? raw : nullptr;
    __syncthreads();           // barrier initialised, counters zeroed

    uint32_t parity = 0;
    for (int sweep = 0;; ++sweep) {
        const int h0 = t0 + sweep * tw;                  // this sweep counts the positions [h0, h0 + tw)
        for (int cb = dx; cb < dy;) {
            mbar_wait(&s_bar, parity);
            parity ^= 1;
            const int n = s_n;
            if (n == 0) break;
            for (int ir = tid; ir < n; ir += MQ_NT) {
                const RInfo r = s_ri[ir];
                const uint32_t cls = r.cls;
                if (cls & CLS_ADMITTED) {
                    if (MINQ && (cls & CLS_COUNTED) && (cls & CLS_HASSEQ)) lowq_prefetch(A, r, cb + ir);
                    if (rawp) {
                        const int a = max(r.pos - h0, 0), b = min(r.end - h0, tw);
                        if (a < b) { atomicAdd(rawp + a, 1u); if (b < tw) atomicAdd(rawp + b, 0u - 1u); }
                    }
                    if (cls & CLS_COUNTED) {
                        const int bank = STRAND ? (int)(cls & 3u) : (int)((cls >> 1) & 1u);
                        const uint32_t q = (cls >> CLS_MAPQ_SHIFT) & 255u;
                        const uint32_t fl = (deep ? 1u : 0u) | (sweep == 0 ? 2u : 0u) | ((uint32_t)bank << 2);
                        // the read has an event at every position of the reference interval [x, x + l)
                        auto emit = [&](int x, int l) { mq_emit<NBK>(cnt, tw, h0, t0, fl, q, x, l); };
                        if (!MINQ && !A.no_del) emit(r.pos, r.end - r.pos);              // M, D and N spans all count (opt.pyx:1904-1906)
                        else if (cls & CLS_NOREF) { if (!A.no_del) emit(r.pos, 1); }     // one deletion-like event at pos
                        else if (!MINQ && (cls & CLS_SIMPLE)) emit(r.pos, r.end - r.pos);
                        else {
                            // no_del: the aligned spans only; min_baseq: of those, the runs of bases at or above the threshold
                            // (a read without sequence has quality 0 everywhere: only its deletion / ref-skip events remain)
                            const bool bases = !MINQ || (cls & CLS_HASSEQ);
                            const uint32_t *mw = MINQ ? A.lowq + (r.seq_off >> 4) + (cb + ir) : nullptr;
                            const uint32_t *cg = A.cigar + r.cigar_off;
                            const int nc = (int)r.n_cigar;
                            int x = r.pos, y = 0;
                            for (int k = 0; k < nc && x < t0 + TT; ++k) {
                                const uint32_t w = __ldg(cg + k);
                                const int op = w & 15, l = (int)(w >> 4);
                                if (op_is_match(op)) {
                                    if (bases && x + l > t0) {
                                        if (MINQ) lowq_gaps(mw, x, l, y, t0, TT, [&](int p, int len) { emit(t0 + p, len); });
                                        else emit(x, l);
                                    }
                                    x += l; y += l;
                                } else if (op == OP_D || op == OP_N) {
                                    if (!A.no_del && x + l > t0) emit(x, l);
                                    x += l;
                                } else if (op == OP_I || op == OP_S) y += l;
                            }
                        }
                    }
                }
            }
            cb += n;
            if (cb >= dy) break;
            __syncthreads();                       // staging buffer free again
            if (tid == 0) issue(cb);
        }
        __syncthreads();
        // ---- the maxima (first sweep): warp b resolves bank b, lane = block ----------------------------------
        if (sweep == 0 && warp < NBK) {
            uint32_t v[MQ_NLEV];
#pragma unroll
            for (int k = 0; k < MQ_NLEV; ++k) v[k] = tab[(k * NBK + warp) * MQ_NB + lane];
#pragma unroll
            for (int k = MQ_NLEV - 1; k >= 1; --k) {       // push-down (imax_pushdown_ref with one lane per block)
                const int h = 1 << (k - 1);
                const uint32_t up = __shfl_up_sync(FULL, v[k], h);
                uint32_t m = max(v[k - 1], v[k]);
                if (lane >= h) m = max(m, up);
                v[k - 1] = m;
            }
            const uint32_t body = v[0];
            // imax_resolve_block on the block's 16 positions, 128-bit accesses
            const int bo = warp * MQ_PS + imax_idx(IMAX_BLK * lane);
            const uint4 *sp = reinterpret_cast<const uint4 *>(sfx + bo);
            const uint4 *pp = reinterpret_cast<const uint4 *>(pfx + bo);
            uint4 *tp = reinterpret_cast<uint4 *>(pt + bo);
            uint32_t s[IMAX_BLK];
            uint32_t run = 0;
#pragma unroll
            for (int j = 0; j < IMAX_BLK / 4; ++j) {
                const uint4 f = sp[j];
                run = max(run, f.x); s[4 * j] = run;
                run = max(run, f.y); s[4 * j + 1] = run;
                run = max(run, f.z); s[4 * j + 2] = run;
                run = max(run, f.w); s[4 * j + 3] = run;
            }
            run = 0;
#pragma unroll
            for (int j = IMAX_BLK / 4 - 1; j >= 0; --j) {
                const uint4 f = pp[j];
                uint4 t = tp[j];
                run = max(run, f.w); t.w = max(max(t.w, body), max(s[4 * j + 3], run));
                run = max(run, f.z); t.z = max(max(t.z, body), max(s[4 * j + 2], run));
                run = max(run, f.y); t.y = max(max(t.y, body), max(s[4 * j + 1], run));
                run = max(run, f.x); t.x = max(max(t.x, body), max(s[4 * j], run));
                tp[j] = t;
            }
        }
        {   // prefix sums of the difference arrays (+ raw); the upper half of the warps first: the lower has the maxima too
            const int nsc = deep ? 4 * NBK : 2 * NBK;
            for (int a = (warp + NBK) & (NWARPS - 1); a < nsc + (A.need_raw ? 1 : 0); a += NWARPS) {
                uint32_t *arr = a < nsc ? cnt + a * tw : raw;
                if (deep) warp_scan_array<TT / 2>(arr, lane);
                else warp_scan_array<TT>(arr, lane);
            }
        }
        __syncthreads();
        // ---- rec(): (group of fields, quad of positions) tasks, 128-bit stores -------------------------------
        {
            const int64_t h0l = t0l + (int64_t)sweep * tw;
            const int qshift = deep ? 6 : 7, nquad = 1 << qshift;                  // tw / 4 quads: a warp holds one group
            const bool aligned = ((reinterpret_cast<uintptr_t>(A.out) & 15) == 0) && ((A.stride & 3) == 0) &&
                                 (((h0l - A.start) & 3) == 0);
            // groups of fields, in the order the tasks are dealt (the root mean squares first: they are the long tasks)
            //   strand: rms[0..2] | rms[3..5] | n[0..5] | nz[0..5] | max[0..5]      else: (rms, max) | (n, nz)
            constexpr int NG = STRAND ? 5 : 2;
#pragma unroll 1
            for (int task = tid; task < NG * nquad; task += MQ_NT) {
                const int g = task >> qshift, qd = task & (nquad - 1);
                const int64_t pl = h0l + 4 * qd;
                const int64_t rem = A.end - pl;
                if (rem <= 0) continue;
                QuadOut<OUT_STORE> O;
                O.p = A.out + (pl - A.start);
                O.stride = A.stride;
                O.seen = 0;
                O.nvalid = rem > 4 ? 4 : (int)rem;
                O.vec = aligned && (O.nvalid == 4 || (pl - A.start) + 4 <= A.stride);
                const uint32_t *cq = cnt + 4 * qd;
                if (STRAND ? g < 2 : g == 0) {
                    // CountPpStrand order: all, fwd, rev, pp, pp_fwd, pp_rev = banks 0123, 02, 13, 23, 2, 3
                    if (STRAND) {
                        O.skip(12 + 3 * g);
                        if (g == 0) { O.put(mq_rms<15, NBK>(cq, tw, deep)); O.put(mq_rms<5, NBK>(cq, tw, deep)); O.put(mq_rms<10, NBK>(cq, tw, deep)); }
                        else { O.put(mq_rms<12, NBK>(cq, tw, deep)); O.put(mq_rms<4, NBK>(cq, tw, deep)); O.put(mq_rms<8, NBK>(cq, tw, deep)); }
                    } else {
                        // two banks: not proper pair, proper pair
                        O.skip(4);
                        O.put(mq_rms<3, NBK>(cq, tw, deep)); O.put(mq_rms<2, NBK>(cq, tw, deep));
                        const uint4 *mp = reinterpret_cast<const uint4 *>(pt + imax_idx(sweep * tw + 4 * qd));
                        const uint4 m_pp = mp[MQ_PS / 4];
                        O.put(max4(mp[0], m_pp));
                        O.put(m_pp);
                    }
                } else if (STRAND && g == 4) {
                    O.skip(18);
                    const uint4 *mp = reinterpret_cast<const uint4 *>(pt + imax_idx(sweep * tw + 4 * qd));
                    const uint4 m0 = mp[0], m1 = mp[MQ_PS / 4], m2 = mp[2 * MQ_PS / 4], m3 = mp[3 * MQ_PS / 4];
                    const uint4 m_pp = max4(m2, m3);
                    O.put(max4(max4(m0, m1), m_pp)); O.put(max4(m0, m2)); O.put(max4(m1, m3)); O.put(m_pp); O.put(m2); O.put(m3);
                } else {
                    // the counts: n (and the row mask), nz
                    const bool zs = STRAND && g == 3;            // this task writes the mapq == 0 counts
                    uint4 c[NBK];
#pragma unroll
                    for (int b = 0; b < NBK; ++b) c[b] = zs ? mq_ldz<NBK>(cq, b, tw, deep) : mq_ldn(cq, b, tw, deep);
                    uint4 c_all = c[0] + c[1];
                    if (STRAND) {
                        c_all = c_all + (c[NBK - 2] + c[NBK - 1]);
                        if (zs) O.skip(6);
                        O.put(c_all); O.put(c[0] + c[NBK - 2]); O.put(c[1] + c[NBK - 1]); O.put(c[NBK - 2] + c[NBK - 1]); O.put(c[NBK - 2]); O.put(c[NBK - 1]);
                    } else {
                        O.put(c_all); O.put(c[1]);
                        const uint4 z1 = mq_ldz<NBK>(cq, 1, tw, deep);
                        O.put(mq_ldz<NBK>(cq, 0, tw, deep) + z1); O.put(z1);
                    }
                    if (!zs) {
                        // the row mask: htslib emits a column wherever an admitted read has an event
                        const uint4 all = A.need_raw ? *reinterpret_cast<const uint4 *>(raw + 4 * qd) : c_all;
                        uint8_t *m = A.rowmask + (pl - A.start);
                        if (O.nvalid == 4 && (reinterpret_cast<uintptr_t>(m) & 3) == 0) {
                            *reinterpret_cast<uint32_t *>(m) = (all.x ? 1u : 0u) | (all.y ? 0x100u : 0u) | (all.z ? 0x10000u : 0u) | (all.w ? 0x1000000u : 0u);
                        } else {
                            if (O.nvalid > 0) m[0] = all.x ? 1 : 0;
                            if (O.nvalid > 1) m[1] = all.y ? 1 : 0;
                            if (O.nvalid > 2) m[2] = all.z ? 1 : 0;
                            if (O.nvalid > 3) m[3] = all.w ? 1 : 0;
                        }
                    }
                }
            }
        }
        if (!deep || sweep == 1) break;
        __syncthreads();           // everyone is done with the counters of the first half
        {
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int k = tid; k < 2 * NBK * TT / 4; k += MQ_NT) reinterpret_cast<uint4 *>(cnt)[k] = z;
            if (A.need_raw) for (int k = tid; k < TT / 4; k += MQ_NT) reinterpret_cast<uint4 *>(raw)[k] = z;
        }
        if (tid == 0) issue(dx);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ tlen, tlen_strand
//
// TlenHelper (opt.pyx:679-727) keeps, per class, n, sum(tlen), sum(tlen^2) -- constant along a read, so pure
// interval adds like the coverage counters (32-bit counts, 64-bit sums, difference arrays + prefix scan)
// -- and Welford's running mean / M2 in doubles, whose rounding depends on the column (file) order.
// The finalisation computes std_tlen = rint(sqrt(M2 / (n - 1))) from the EXACT integer sums:
// M2 = A / n with A = n sum(t^2) - sum(t)^2, so the decision between k and k + 1 is the sign of
// 4 A - (2k + 1)^2 n (n - 1), evaluated in 128-bit integers.  Welford's computed M2 differs from the exact
// one by at most a few n u kappa^2 relative (u = 2^-53, kappa^2 = n sum(t^2) / A the condition number of the
// variance; Chan, Golub & LeVeque 1983), so whenever |4 A - (2k + 1)^2 n (n - 1)| is above
// 256 n^2 sum(t^2) u -- 64x that bound -- the rounded value is the reference's whatever the order.
// Positions that are closer to a rounding boundary than that (in practice: exact ties, e.g. std = 0.5)
// put their tile on the replay list, and k_pileup<AccTlen> -- the in-order Welford replay -- recomputes the listed
// tiles right after this kernel (a small fixed grid walking the list).
//
// word arrays of TT.  tlen_strand: nfwd[0] nrev[1] | class c = paired fwd, paired rev, pp fwd, pp rev: n[2 + c] (32-bit
// difference arrays) | sum[c] 64-bit at words 6.., sumsq[c] 64-bit at words 14.. (64-bit difference arrays): 22 words per
// position, 256 threads x 4 CTAs per SM.  tlen reports all / paired / pp without the strand split: nall[0] | class
// c = paired, pp: n[1 + c] | sum[c] at words 3.., sumsq[c] at words 7..: 11 words, half the arrays to zero and scan,
// 128 threads x 8 CTAs per SM (27 KB each) -- the layout that took mapq from 0.82 to 0.53 ms per 50 Mb.
template <int STAT> struct TlCfg { static const int NN = 2, NC = 4, NT = 256, CTAS = 4, CH = 160; };       // tlen_strand
template <> struct TlCfg<PSS_TLEN> { static const int NN = 1, NC = 2, NT = 128, CTAS = 8, CH = 136; };
static_assert(PSS_TTILE == 2 * PSS_TILE, "k_tilet hands the two halves of its tile to the replay kernel");

// 64-bit add into shared memory as two 32-bit atomic adds: the carry out of the low word is seen by exactly the add that
// causes it (old + lo wraps), so low and high word together always hold the sum of all increments mod 2^64 whatever the
// interleaving.  The native form is a compare-and-swap loop (ATOMS.CAST.SPIN.64: a load, the CAS, retries under
// contention); a typical increment here -- a template length, or its two's complement at the interval's end, whose high
// word 0xffffffff is cancelled by the carry -- is ONE 32-bit add.
__device__ __forceinline__ void sadd64(unsigned long long *p, unsigned long long inc) {
    uint32_t *w = reinterpret_cast<uint32_t *>(p);
    const uint32_t lo = (uint32_t)inc;
    uint32_t hi = (uint32_t)(inc >> 32);
    if (lo) {
        const uint32_t old = atomicAdd(w, lo);
        hi += ((uint32_t)(old + lo) < lo) ? 1u : 0u;
    }
    if (hi) atomicAdd(w + 1, hi);
}
__device__ __forceinline__ void interval_add64(unsigned long long *arr, int x, int l, int t0, unsigned long long inc) {
    int a = max(x - t0, 0), b = min(x + l - t0, TT);
    if (a < b) {
        sadd64(arr + a, inc);
        if (b < TT) sadd64(arr + b, 0ull - inc);
    }
}

// in-place inclusive prefix sum of TT 64-bit words by one warp (lane l owns words [128 k + 4 l, 128 k + 4 l + 4): four
// rounds of the five-step shuffle scan instead of eight with two words per lane)
__device__ __forceinline__ void warp_scan_array64(unsigned long long *arr, int lane) {
    unsigned long long carry = 0;
#pragma unroll 1
    for (int k = 0; k < TT / 128; ++k) {
        ulonglong2 *p = reinterpret_cast<ulonglong2 *>(arr + 128 * k + 4 * lane);
        ulonglong2 u = p[0], v = p[1];
        u.y += u.x; v.x += u.y; v.y += v.x;
        unsigned long long inc = v.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const unsigned long long off = carry + inc - v.y;
        carry += __shfl_sync(FULL, inc, 31);
        u.x += off; u.y += off; v.x += off; v.y += off;
        p[0] = u; p[1] = v;
    }
}

// std from the exact sums; *amb: the reference's (order dependent) Welford value could round differently
template <class U>       // U = unsigned long long when every product below fits 64 bits, else unsigned __int128
__device__ __forceinline__ int tlen_std_t(unsigned long long n, U ss, unsigned long long Q, bool *amb, float force) {
    const U nq = (U)n * (U)Q;
    const U A = nq >= ss ? nq - ss : 0;                    // n sum(t^2) - sum(t)^2 >= 0
    const U B = (U)n * (U)(n - 1);
    const double est = sqrt((double)A / (double)B);
    unsigned long long k = est > 0.5 ? (unsigned long long)(est - 0.5) : 0;        // <= floor(sqrt(A / B)) up to rounding
    while ((U)k * k * B > A) --k;
    while ((U)(k + 1) * (k + 1) * B <= A) ++k;
    const U T = (U)(2 * k + 1) * (2 * k + 1) * B, F = 4 * A;
    const U D = F > T ? F - T : T - F;
    const double thr = 256.0 * (double)n * (double)n * (double)Q * 1.1102230246251565e-16 + force;
    if (D == 0 || (double)D <= thr) *amb = true;
    return (int)(k + (F > T ? 1u : F == T ? (k & 1u) : 0u));
}
__device__ __forceinline__ int tlen_std(long long n, long long S, unsigned long long Q, bool *amb, float force) {
    if (n < 2) return 0;
    // n sum(t^2) < 2^60: sum(t)^2 <= n sum(t^2), 4 A and (2k + 1)^2 n (n - 1) ~ 4 A stay below 2^63
    if (n < (1 << 20) && Q < (1ull << 40)) {
        const unsigned long long a = (unsigned long long)(S < 0 ? -S : S);
        return tlen_std_t<unsigned long long>((unsigned long long)n, a * a, Q, amb, force);
    }
    return tlen_std_t<unsigned __int128>((unsigned long long)n, (unsigned __int128)((__int128)S * (__int128)S), Q, amb, force);
}

struct TlenOut { int n, mean, rms, sd; };

// The exact forms, out of line: what the finalisation used for every class of every position until ncu showed them to be
// ~80 % of tlen_strand's instructions (a double division + square root + 64- / 128-bit boundary products per value, inlined
// six times: 4976 SASS instructions).  Now only the values the float screening below cannot decide come here.
__device__ __noinline__ int tlen_mean_exact(long long S, unsigned n) {
    return n > 0 ? (int)rint((double)S / (double)n) : 0;                    // opt.pyx:2166-2170
}
__device__ __forceinline__ int tlen_rms_exact(unsigned long long Q, unsigned n) { return q_rootmean_exact(Q, n); }   // opt.pyx:2159-2163
__device__ __noinline__ unsigned tlen_std_exact(unsigned n, long long S, unsigned long long Q, float force) {   // sd | ambiguous << 31
    bool amb = false;
    const int sd = tlen_std((long long)n, S, Q, &amb, force);
    return (unsigned)sd | (amb ? 0x80000000u : 0u);
}

// mean, rms and std of one class from n, sum(t), sum(t^2).  Each is rint() of a real number; a FLOAT estimate e of that
// number is off by less than 5e-7 |e| (conversions 2^-24 each, the approximate reciprocal / division and reciprocal square
// root 2 ulp each, the products 2^-24; the square root halves what enters under it), so whenever
// |e - rint(e)| <= 0.4999 - 1e-6 |e| the real number is less than 0.5 away from rint(e): that IS its rounding and no tie
// is near.  For std the margin is 0.4995: the order-dependent Welford value of the reference can differ from the exact one
// only within the ambiguity bound of tlen_std_t, which with sum(t^2) < 2^32 is at most 2.4e-4 in std -- so an accepted
// position need not be replayed either.  Everything else -- estimates next to a half (one value in a thousand), sums beyond
// 32 bits, n >= 65536, the forced replay of the tests -- takes the exact forms above.  scripts/tlen_fast_check.cpp runs
// this screening against the exact forms on adversarial and random multisets with the estimates perturbed by 1e-6.
__device__ __forceinline__ TlenOut tlen_class(unsigned n, long long S, unsigned long long Q, bool *amb, float force) {
    TlenOut o;
    o.n = (int)n;
    if (n == 0) { o.mean = o.rms = o.sd = 0; return o; }
    const bool small = force == 0.0f && n < 65536u && (Q >> 32) == 0 && (long long)(int)S == S;
    const float inv = __fdividef(1.0f, (float)n);
    {   // mean
        const float e = (float)(int)S * inv, r = rintf(e);
        o.mean = (small && fabsf(e - r) <= 0.4999f - 1e-6f * fabsf(e)) ? (int)r : tlen_mean_exact(S, n);
    }
    {   // root mean square
        const float x = (float)(unsigned)Q * inv, e = x * rsqrtf(x), r = rintf(e);      // (Q == 0: NaN, never accepted)
        o.rms = Q == 0 ? 0 : (small && fabsf(e - r) <= 0.4999f - 1e-6f * e) ? (int)r : tlen_rms_exact(Q, n);
    }
    if (n < 2) { o.sd = 0; return o; }
    {   // standard deviation: sqrt(A / (n (n - 1))), A = n sum(t^2) - sum(t)^2 exactly (n < 2^16, sum(t^2) < 2^32, |sum(t)| < 2^31)
        const unsigned long long nq = (unsigned long long)n * (unsigned)Q, ss = (unsigned long long)(S * S);
        const unsigned long long Aex = nq >= ss ? nq - ss : 0;
        const float v = __fdividef((float)Aex, (float)(n * (n - 1u))), e = v * rsqrtf(v), r = rintf(e);   // (A == 0: NaN)
        if (small && (Aex == 0 || fabsf(e - r) <= 0.4995f - 1e-6f * e)) o.sd = Aex == 0 ? 0 : (int)r;
        else {
            const unsigned w = tlen_std_exact(n, S, Q, force);
            o.sd = (int)(w & 0x7fffffffu);
            if (w >> 31) *amb = true;
        }
    }
    return o;
}

template <int STAT>
__global__ void __launch_bounds__(TlCfg<STAT>::NT, TlCfg<STAT>::CTAS)
k_tilet(TileArgs A, const int32_t *__restrict__ tlen, int32_t *__restrict__ redo, float force) {
    using C = TlCfg<STAT>;
    constexpr bool STRAND = STAT == PSS_TLEN_STRAND;
    constexpr int NN = C::NN, NC = C::NC, TLEN_NW = NN + 5 * NC, TLEN_NT = C::NT, TCHK = C::CH;
    constexpr int NWARPS = TLEN_NT / 32;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);
    uint32_t *raw = cnt + TLEN_NW * TT;                                   // exists only when need_raw
    RInfo *s_ri = reinterpret_cast<RInfo *>(raw + (A.need_raw ? TT : 0));
    int32_t *s_tlen = reinterpret_cast<int32_t *>(s_ri + TCHK);           // TCHK + 8
    unsigned long long *sum = reinterpret_cast<unsigned long long *>(cnt + (NN + NC) * TT);
    unsigned long long *sumsq = sum + NC * TT;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_n;
    __shared__ long long s_tlenlo;
    __shared__ int s_amb;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t t0l = A.start + (int64_t)tile * TT;
    const int t0 = (int)t0l;
    int dx, dy;
    {
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(A.desc + tile));
        dx = (int)d0.x; dy = (int)d0.y;
    }
    auto issue = [&](int cb) {
        const int n = min(TCHK, dy - cb);
        uint32_t b1 = 0;
        const uint8_t *p1 = nullptr;
        if (n > 0) stage_span(tlen, (long long)cb, (long long)cb + n, 4, &p1, &b1, &s_tlenlo);
        s_n = n;
        fence_proxy_async();
        if (n > 0) {
            mbar_arrive_expect_tx(&s_bar, (uint32_t)n * (uint32_t)sizeof(RInfo) + b1);
            bulk_g2s(s_ri, A.ri + cb, (uint32_t)n * (uint32_t)sizeof(RInfo), &s_bar);
            bulk_g2s(s_tlen, p1, b1, &s_bar);
        } else mbar_arrive_expect_tx(&s_bar, 0);
    };
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        s_amb = 0;
        issue(dx);
    }
#if TPREFETCH > 0
    if (tid == TLEN_NT - 32 && tile + TPREFETCH < (int)gridDim.x) {
        const uint4 p0 = __ldg(reinterpret_cast<const uint4 *>(A.desc + tile + TPREFETCH));
        const int px = (int)p0.x, pn = min(TCHK, (int)p0.y - px);
        if (pn > 0) bulk_prefetch_l2(A.ri + px, (uint32_t)pn * (uint32_t)sizeof(RInfo));
    }
#endif
    {
        const int nw4 = (TLEN_NW + (A.need_raw ? 1 : 0)) * TT / 4;
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < nw4; k += TLEN_NT) reinterpret_cast<uint4 *>(cnt)[k] = z;
    }
    uint32_t *rawp = A.need_raw ? raw : nullptr;
    const bool minq = A.lowq != nullptr;
    __syncthreads();
    uint32_t parity = 0;
    for (int cb = dx; cb < dy;) {
        mbar_wait(&s_bar, parity);
        parity ^= 1;
        const int n = s_n;
        if (n == 0) break;
        const int tskew = (int)(cb - s_tlenlo);
        for (int ir = tid; ir < n; ir += TLEN_NT) {
            const RInfo r = s_ri[ir];
            const uint32_t cls = r.cls;
            if (cls & CLS_ADMITTED) {
                const int len = r.end - r.pos;                    // every event of the column counts (opt.pyx:700-727)
                if (rawp) interval_add(rawp, r.pos, len, t0, 1u);
                if (cls & CLS_COUNTED) {
                    const int rev = (STRAND && (cls & CLS_REV)) ? 1 : 0;
                    const long long t = s_tlen[ir + tskew];
                    // the read's contribution to [x, x + l): +1 / +t / +t^2 per class (sign: taken back out again)
                    auto contrib = [&](int x, int l, bool neg) {
                        interval_add(cnt + rev * TT, x, l, t0, neg ? 0u - 1u : 1u);
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            if (cls & (g ? CLS_PP : CLS_PAIRED)) {
                                const int c = STRAND ? 2 * g + rev : g;
                                interval_add(cnt + (NN + c) * TT, x, l, t0, neg ? 0u - 1u : 1u);
                                if (t) {
                                    const unsigned long long ut = (unsigned long long)t, uq = (unsigned long long)(t * t);
                                    interval_add64(sum + c * TT, x, l, t0, neg ? 0ull - ut : ut);
                                    interval_add64(sumsq + c * TT, x, l, t0, neg ? 0ull - uq : uq);
                                }
                            }
                        }
                    };
                    if (A.no_del) {
                        // deletion-like events are dropped (opt.pyx:1904-1906): the aligned spans only, and under min_baseq
                        // the runs of bases at or above the threshold in them (none for a read without sequence)
                        if (!(cls & CLS_NOREF) && (!minq || (cls & CLS_HASSEQ))) {
                            if (minq) lowq_prefetch(A, r, cb + ir);
                            if (!minq && (cls & CLS_SIMPLE)) contrib(r.pos, len, false);
                            else {
                                const uint32_t *mw = minq ? A.lowq + (r.seq_off >> 4) + (cb + ir) : nullptr;
                                const uint32_t *cg = A.cigar + r.cigar_off;
                                int x = r.pos, y = 0;
                                for (int k = 0; k < (int)r.n_cigar && x < t0 + TT; ++k) {
                                    const uint32_t w = __ldg(cg + k);
                                    const int op = w & 15, l = (int)(w >> 4);
                                    if (op_is_match(op)) {
                                        if (x + l > t0) {
                                            if (minq) lowq_gaps(mw, x, l, y, t0, TT, [&](int p, int run) { contrib(t0 + p, run, false); });
                                            else contrib(x, l, false);
                                        }
                                        x += l; y += l;
                                    } else if (op == OP_D || op == OP_N) x += l;
                                    else if (op == OP_I || op == OP_S) y += l;
                                }
                            }
                        }
                    } else if (!minq) contrib(r.pos, len, false);
                    else if (!(cls & CLS_HASSEQ)) {
                        // min_baseq > 0, no sequence: quality 0 at every base -- only the deletion / ref-skip events remain
                        if (cls & CLS_NOREF) contrib(r.pos, 1, false);
                        else {
                            const uint32_t *cg = A.cigar + r.cigar_off;
                            int x = r.pos;
                            for (int k = 0; k < (int)r.n_cigar && x < t0 + TT; ++k) {
                                const uint32_t w = __ldg(cg + k);
                                const int op = w & 15, l = (int)(w >> 4);
                                if (op_is_match(op)) x += l;
                                else if (op == OP_D || op == OP_N) { if (x + l > t0) contrib(x, l, false); x += l; }
                            }
                        }
                    } else {
                        // min_baseq > 0: as if every base passed, then the marked bases (k_lowq) are taken back out
                        lowq_prefetch(A, r, cb + ir);
                        contrib(r.pos, len, false);
                        if (!(cls & CLS_NOREF)) {
                            const uint32_t *mw = A.lowq + (r.seq_off >> 4) + (cb + ir);
                            auto undo = [&](int p, int run) { contrib(t0 + p, run, true); };
                            if (cls & CLS_SIMPLE) lowq_runs(mw, r.pos, len, 0, t0, undo);
                            else {
                                const uint32_t *cg = A.cigar + r.cigar_off;
                                int x = r.pos, y = 0;
                                for (int k = 0; k < (int)r.n_cigar && x < t0 + TT; ++k) {
                                    const uint32_t w = __ldg(cg + k);
                                    const int op = w & 15, l = (int)(w >> 4);
                                    if (op_is_match(op)) { if (x + l > t0) lowq_runs(mw, x, l, y, t0, undo); x += l; y += l; }
                                    else if (op == OP_D || op == OP_N) x += l;
                                    else if (op == OP_I || op == OP_S) y += l;
                                }
                            }
                        }
                    }
                }
            }
        }
        cb += n;
        if (cb >= dy) break;
        __syncthreads();
        if (tid == 0) issue(cb);
    }
    __syncthreads();
    {   // NN + NC (+ raw) 32-bit and 2 NC 64-bit difference arrays; the 64-bit ones (the longer scans) first
        const int n64 = 2 * NC, n32 = NN + NC;
        for (int a = warp; a < n64 + n32 + (A.need_raw ? 1 : 0); a += NWARPS) {
            if (a < n64) warp_scan_array64(sum + a * TT, lane);            // sum[] and sumsq[] are contiguous
            else if (a < n64 + n32) warp_scan_array(cnt + (a - n64) * TT, lane);
            else warp_scan_array(raw, lane);
        }
    }
    __syncthreads();
    const int64_t st = A.stride;
    bool amb = false;
#pragma unroll 1
    for (int p = tid; p < TT; p += TLEN_NT) {
        const int64_t pl = t0l + p;
        if (pl >= A.end) break;
        int32_t *o = A.out + (pl - A.start);
        const uint32_t nf = cnt[p], nr = STRAND ? cnt[TT + p] : 0u;       // (no strand split: nf counts every read)
        uint32_t nc[NC];
        long long S[NC];
        unsigned long long Q[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) { nc[c] = cnt[(NN + c) * TT + p]; S[c] = (long long)sum[c * TT + p]; Q[c] = sumsq[c * TT + p]; }
        if (!STRAND) {
            const TlenOut h0 = tlen_class(nc[0], S[0], Q[0], &amb, force);
            const TlenOut h1 = tlen_class(nc[1], S[1], Q[1], &amb, force);
            o[0] = nf; o[st] = h0.n; o[2 * st] = h1.n; o[3 * st] = h0.mean; o[4 * st] = h1.mean;
            o[5 * st] = h0.rms; o[6 * st] = h1.rms; o[7 * st] = h0.sd; o[8 * st] = h1.sd;
        } else {
            o[0] = nf + nr; o[st] = nf; o[2 * st] = nr;
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {          // paired, pp: all, fwd, rev
                const int c = (2 * g) & (NC - 1), c1 = (c + 1) & (NC - 1);        // (masks: NC == 4 on this path)
                const TlenOut h[3] = {tlen_class(nc[c] + nc[c1], S[c] + S[c1], Q[c] + Q[c1], &amb, force),
                                      tlen_class(nc[c], S[c], Q[c], &amb, force), tlen_class(nc[c1], S[c1], Q[c1], &amb, force)};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int f = 3 * g + k;
                    o[(3 + f) * st] = h[k].n; o[(9 + f) * st] = h[k].mean; o[(15 + f) * st] = h[k].rms; o[(21 + f) * st] = h[k].sd;
                }
            }
        }
        A.rowmask[pl - A.start] = (A.need_raw ? raw[p] : nf + nr) ? 1 : 0;
    }
    if (amb) s_amb = 1;
    __syncthreads();
    if (tid == 0 && s_amb) {
        // both 256-position halves of this tile go on the replay list (redo[0] = its length)
        const int g0 = 2 * tile, gtiles = (int)((A.end - A.start + PSS_TILE - 1) / PSS_TILE);
        const int two = g0 + 1 < gtiles ? 1 : 0;
        const int k = atomicAdd(redo, 1 + two);
        redo[1 + k] = g0;
        if (two) redo[2 + k] = g0 + 1;
    }
}

// ------------------------------------------------------- descriptors + packed reference

// One thread per tile: candidate read range (reads are start-sorted, so it is the index
// range [first pos >= t0 - maxspan + 1, first pos >= t1)), the first chunk and its sequence span.
__global__ void __launch_bounds__(256)
k_tile_desc(const RInfo *__restrict__ ri, const int64_t *__restrict__ seq_off, int64_t n, SmallState *st,
            int64_t start, int64_t ntiles, int with_seq, TileDesc *__restrict__ desc) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    long long t0 = start + t * TT, t1 = t0 + TT;
    long long key0 = t0 - st->maxspan + 1;
    // first index with pos >= key, searched outward from a guess (exponential steps, then bisection): with
    // evenly spread reads that is a handful of probes instead of log2(n) dependent loads
    auto lower = [&](long long key, int64_t guess) -> int64_t {
        if (n <= 0) return 0;
        if (guess < 0) guess = 0;
        if (guess > n - 1) guess = n - 1;
        int64_t lo, hi;                       // invariant at the end: pos[lo - 1] < key (or lo == 0), pos[hi] >= key (or hi == n)
        if ((long long)ri[guess].pos < key) {
            lo = guess + 1; hi = n;
            for (int64_t step = 32, p = guess + step; ; step <<= 1, p += step) {
                if (p >= n) break;
                if ((long long)ri[p].pos < key) lo = p + 1; else { hi = p; break; }
            }
        } else {
            hi = guess; lo = 0;
            for (int64_t step = 32, p = guess - step; ; step <<= 1, p -= step) {
                if (p < 0) break;
                if ((long long)ri[p].pos < key) { lo = p + 1; break; } else hi = p;
            }
        }
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((long long)ri[mid].pos < key) lo = mid + 1; else hi = mid; }
        return lo;
    };
    const long long p_first = n > 0 ? ri[0].pos : 0, p_last = n > 0 ? ri[n - 1].pos : 0;
    const double dens = p_last > p_first ? (double)(n - 1) / (double)(p_last - p_first) : 0.0;
    const int64_t x = lower(key0, (int64_t)((double)(key0 - p_first) * dens));
    const int64_t y = lower(t1, x + (int64_t)((double)(t1 - key0) * dens));
    TileDesc d;
    d.x = (int32_t)x; d.y = (int32_t)y; d.flags = 0; d.seq_lo = 0; d.seq_hi = 0; d.cig_lo = 0; d.cig_hi = 0;
    int m = (int)(y - x < TCH ? y - x : TCH);
    if (with_seq && m > 0) {
        long long slo = seq_off[x], shi = seq_off[x + m];
        while (m > 1 && shi - slo + 32 > TSEQ_BYTES) { m >>= 1; shi = seq_off[x + m]; }
        if (shi - slo + 32 > TSEQ_BYTES) { atomicOr(&st->toolong, 1); m = 0; }
        d.seq_lo = slo; d.seq_hi = shi;
        if (m > 0) {
            long long clo = ri[x].cigar_off, chi = ri[x + m - 1].cigar_off + ri[x + m - 1].n_cigar;
            if (chi - clo + 8 <= TCIG_WORDS) { d.cig_lo = clo; d.cig_hi = chi; }
        }
    }
    d.n0 = m;
    desc[t] = d;
}

// The reference as 4-bit codes, eight positions per word (first position in the top nibble),
// tile-aligned to `start`; flags tiles that contain a position without a code.
__global__ void __launch_bounds__(256)
k_refpack(const uint8_t *__restrict__ ref, int64_t ref_start, int64_t ref_len, int64_t start, int64_t end,
          int64_t nwords, uint32_t *__restrict__ out, uint32_t *tflags) {
    // the code table in shared memory: lanes look up different characters, which the constant
    // cache would serialise
    __shared__ uint8_t s_code[256];
    s_code[threadIdx.x] = c_refcode[threadIdx.x];
    __syncthreads();
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    int64_t p0 = start + w * 8;
    uint32_t v = 0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int64_t p = p0 + k, ro = p - ref_start;
        int c = (ref && ro >= 0 && ro < ref_len) ? s_code[__ldg(ref + ro)] : 16;
        if (c == 16 && p < end) bad = true;
        v |= (uint32_t)(c & 15) << (28 - 4 * k);
    }
    out[w] = v;
    if (bad) atomicOr(&tflags[(w * 8) / TT], 1u);
}

template <int KIND>
static size_t tile_smem_bytes(bool need_raw = true) {
    size_t narr = KindInfo<KIND>::NDIFF + KindInfo<KIND>::NPOINT + (need_raw ? 1 : 0);   // + raw (the row mask's own counter)
    size_t b = narr * TT * 4 + sizeof(RInfo) * ((KIND == KIND_VAR || KIND == KIND_VAR1) ? TCH : TCH_COV);
    if (KIND == KIND_VAR || KIND == KIND_VAR1) b += (REFN_WORDS + 4) * 4 + TSEQ_BYTES + 32 + (TCIG_WORDS + 8) * 4 + (TBAL ? SPAN_MAX * sizeof(SpanRec) : 0) + (LayVar::DEFER ? EV_SLOTS * EV_STRIDE * 2 : 0);
    return b;
}

template <int KIND, int STAT, bool MINQ>
static cudaError_t launch_tile_mq(const TileArgs &a, int64_t ntiles, cudaStream_t s) {
    size_t smem = tile_smem_bytes<KIND>();
    static bool configured_dev[64] = {false};     // per instantiation and device (attributes are per device)
    int dev__ = 0;
    cudaGetDevice(&dev__);
    bool &configured = configured_dev[dev__ & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tile<KIND, STAT, MINQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_tile<KIND, STAT, MINQ><<<(unsigned)ntiles, KindLaunch<KIND>::NT, tile_smem_bytes<KIND>(a.need_raw != 0), s>>>(a);
    return cudaGetLastError();
}
template <int KIND, int STAT>
static cudaError_t launch_tile_one(const TileArgs &a, int64_t ntiles, cudaStream_t s) {
    return a.lowq ? launch_tile_mq<KIND, STAT, true>(a, ntiles, s) : launch_tile_mq<KIND, STAT, false>(a, ntiles, s);
}

// one bit per read base: quality (after the overlap rewrite) below min_baseq.  One warp per read; the bits of
// read i start at word (seq_off[i] >> 4) + i (its packed sequence takes ceil(l / 2) bytes, so that leaves at
// least ceil(l / 32) words per read without a second offset array).
__global__ void __launch_bounds__(256)
k_lowq(const uint8_t *__restrict__ qual, const int64_t *__restrict__ qual_off, const int64_t *__restrict__ seq_off,
       const int32_t *__restrict__ l_seq, int64_t n, int min_baseq, uint32_t *__restrict__ out) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int l = l_seq[i];
    const uint8_t *q = qual + qual_off[i];
    uint32_t *o = out + (seq_off[i] >> 4) + i;
    for (int b = 0; b < l; b += 32) {
        const bool low = (b + lane < l) && (int)q[b + lane] < min_baseq;
        const uint32_t m = __ballot_sync(FULL, low);
        if (lane == 0) o[b >> 5] = m;
    }
}

size_t tile_lowq_bytes(int64_t n_reads, int64_t n_seq_bytes) { return sizeof(uint32_t) * (size_t)((n_seq_bytes >> 4) + n_reads + 8); }

cudaError_t launch_lowq(const pss_reads &R, const uint8_t *qual, int min_baseq, void *lowq_buf, cudaStream_t s, int64_t *launches) {
    if (R.n_reads <= 0) return cudaSuccess;
    k_lowq<<<(unsigned)((R.n_reads * 32 + 255) / 256), 256, 0, s>>>(qual, R.qual_off, R.seq_off, R.l_seq, R.n_reads, min_baseq,
                                                                    (uint32_t *)lowq_buf);
    if (launches) ++*launches;
    return cudaGetLastError();
}

template <int KIND, int STAT, bool MINQ = false>
static cudaError_t launch_tileq_one(const TileArgs &a, int64_t ntiles, cudaStream_t s) {
    const size_t smem = (size_t)q_layout<KIND, QBanks<KIND, STAT>::NBK>().total;
    static bool configured_dev[64] = {false};     // per instantiation and device (attributes are per device)
    int dev__ = 0;
    cudaGetDevice(&dev__);
    bool &configured = configured_dev[dev__ & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tileq<KIND, STAT, MINQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_tileq<KIND, STAT, MINQ><<<(unsigned)ntiles, QKind<KIND>::NT, smem, s>>>(a);
    return cudaGetLastError();
}

template <int STAT>
static cudaError_t launch_tilem_one(const TileArgs &a, int64_t ntiles, cudaStream_t s) {
    const size_t smem_max = (size_t)(mq_words(MqCfg<STAT>::NBK) + TT) * 4 + (size_t)MqCfg<STAT>::CH * sizeof(RInfo);
    static bool configured_dev[64] = {false};     // per instantiation and device (attributes are per device)
    int dev__ = 0;
    cudaGetDevice(&dev__);
    bool &configured = configured_dev[dev__ & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tilem<STAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tilem<STAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const size_t smem = smem_max - (a.need_raw ? 0 : (size_t)TT * 4);
    if (a.lowq) k_tilem<STAT, true><<<(unsigned)ntiles, MqCfg<STAT>::NT, smem, s>>>(a);
    else k_tilem<STAT, false><<<(unsigned)ntiles, MqCfg<STAT>::NT, smem, s>>>(a);
    return cudaGetLastError();
}

// statistics whose counts come back through the narrow (8- / 16-bit) host transport
bool tile_kernel_counting(int stat) {
    switch (stat) {
    case PSS_COVERAGE: case PSS_COVERAGE_STRAND: case PSS_COVERAGE_GC: case PSS_COVERAGE_EXT:
    case PSS_COVERAGE_EXT_STRAND: case PSS_VARIATION: case PSS_VARIATION_STRAND: return true;
    default: return false;
    }
}

size_t tile_redo_bytes(int64_t ntiles) { return sizeof(int32_t) * (size_t)(2 * ntiles + 2); }    // length + two halves per tile

template <int STAT>
static cudaError_t launch_tilet_one(const TileArgs &a, const int32_t *tlen, int32_t *redo, int64_t ntiles, cudaStream_t s) {
    using C = TlCfg<STAT>;
    const size_t smem_max = (size_t)(C::NN + 5 * C::NC + 1) * TT * 4 + (size_t)C::CH * sizeof(RInfo) + (C::CH + 8) * 4;
    const size_t smem = smem_max - (a.need_raw ? 0 : (size_t)TT * 4);
    static bool configured_dev[64] = {false};     // per instantiation and device (attributes are per device)
    int dev__ = 0;
    cudaGetDevice(&dev__);
    bool &configured = configured_dev[dev__ & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tilet<STAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const float force = getenv("PSS_TLEN_FORCE_REPLAY") ? 3.0e38f : 0.0f;   // tests: every tile takes the replay
    k_tilet<STAT><<<(unsigned)ntiles, C::NT, smem, s>>>(a, tlen, redo, force);
    return cudaGetLastError();
}

bool tile_kernel_supports(int stat, int min_baseq, int no_del) {
    if (tile_kernel_counting(stat)) return true;      // min_baseq > 0: the bases below it are taken back out (k_lowq)
    switch (stat) {
    case PSS_TLEN: case PSS_TLEN_STRAND: return true;        // intervals of events: whole reads, aligned spans under no_del, kept runs under min_baseq
    case PSS_MAPQ: case PSS_MAPQ_STRAND: return true;        // k_tilem: intervals of events, whatever the filters
    case PSS_BASEQ: case PSS_BASEQ_STRAND: return true;      // min_baseq: k_lowq marks for the counts, the threshold in the gather
    case PSS_BASEQ_EXT: case PSS_BASEQ_EXT_STRAND: return min_baseq <= 0;
    default: return false;
    }
}

size_t tile_desc_bytes(int64_t ntiles) { return sizeof(TileDesc) * (size_t)ntiles; }
size_t tile_refpack_bytes(int64_t ntiles) { return ((size_t)ntiles * REFN_WORDS + 8 + (size_t)ntiles) * 4; }   // packed reference + tile flags

cudaError_t launch_tile(int stat, const PileupArgs &p, const int64_t *seq_off, int64_t n_reads, SmallState *st,
                        int need_raw, int64_t ntiles, void *desc_buf, void *refpack_buf, bool desc_valid,
                        cudaStream_t s, int64_t *launches, cudaEvent_t e0, cudaEvent_t e1) {
    if (ntiles <= 0) return cudaSuccess;
    static bool table_ready_dev[64] = {false};   // the __constant__ table exists per device
    int dev_t = 0;
    cudaGetDevice(&dev_t);
    bool &table_ready = table_ready_dev[dev_t & 63];
    if (!table_ready) {
        uint8_t t[256];
        refcode_table(t);
        cudaError_t e = cudaMemcpyToSymbol(c_refcode, t, 256);
        if (e != cudaSuccess) return e;
        table_ready = true;
    }
    const bool var = (stat == PSS_VARIATION || stat == PSS_VARIATION_STRAND || stat == PSS_BASEQ_EXT ||
                      stat == PSS_BASEQ_EXT_STRAND);
    TileDesc *desc = (TileDesc *)desc_buf;
    if (!desc_valid) {
        // always with the sequence / CIGAR spans: the descriptors are shared by all statistics of the batch
        k_tile_desc<<<(unsigned)((ntiles + 255) / 256), 256, 0, s>>>(p.ri, seq_off, n_reads, st, p.start, ntiles, 1, desc);
        if (launches) ++*launches;
    }
    if (var) {
        int64_t nwords = ntiles * REFN_WORDS + 8;
        uint32_t *tflags = (uint32_t *)refpack_buf + nwords;
        cudaError_t em = cudaMemsetAsync(tflags, 0, sizeof(uint32_t) * (size_t)ntiles, s);
        if (em != cudaSuccess) return em;
        k_refpack<<<(unsigned)((nwords + 255) / 256), 256, 0, s>>>(p.ref, p.ref_start, p.ref_len, p.start, p.end, nwords,
                                                                  (uint32_t *)refpack_buf, tflags);
        if (launches) ++*launches;
    }
    TileArgs a;
    a.ri = p.ri; a.cigar = p.cigar; a.seq4 = p.seq4; a.seq_off = seq_off; a.desc = desc;
    a.qual = p.qual; a.qual_off = p.qual_off;
    a.lowq = p.min_baseq > 0 ? p.lowq : nullptr;
    a.min_baseq = p.min_baseq;
    a.refpack = (const uint32_t *)refpack_buf;
    a.tflags = (const uint32_t *)refpack_buf + (ntiles * REFN_WORDS + 8);
    a.ref = p.ref; a.ref_start = p.ref_start; a.ref_len = p.ref_len; a.start = p.start; a.end = p.end;
    a.stride = p.out_stride;
    a.no_del = p.no_del; a.need_raw = need_raw; a.out = p.out; a.rowmask = p.rowmask; a.st = st;
    a.narrow = p.narrow; a.ovf8 = p.ovf8;
    a.timing = nullptr;
#ifdef TILE_TIMING
    static unsigned long long *tbuf = nullptr;
    static size_t tcap = 0;
    size_t need = ((size_t)ntiles / 64 + 1) * 8;
    if (need > tcap) { if (tbuf) cudaFree(tbuf); cudaMalloc(&tbuf, need * 8); tcap = need; }
    cudaMemsetAsync(tbuf, 0, need * 8, s);
    a.timing = tbuf;
#endif
    if (launches) ++*launches;
    if (e0) cudaEventRecord(e0, s);
    cudaError_t e;
    switch (stat) {
    case PSS_COVERAGE: e = launch_tile_one<KIND_COV, PSS_COVERAGE>(a, ntiles, s); break;
    case PSS_COVERAGE_STRAND: e = launch_tile_one<KIND_COV, PSS_COVERAGE_STRAND>(a, ntiles, s); break;
    case PSS_COVERAGE_GC: e = launch_tile_one<KIND_COV, PSS_COVERAGE_GC>(a, ntiles, s); break;
    case PSS_COVERAGE_EXT: e = launch_tile_one<KIND_EXT, PSS_COVERAGE_EXT>(a, ntiles, s); break;
    case PSS_COVERAGE_EXT_STRAND: e = launch_tile_one<KIND_EXT, PSS_COVERAGE_EXT_STRAND>(a, ntiles, s); break;
    case PSS_VARIATION: e = launch_tile_one<KIND_VAR1, PSS_VARIATION>(a, ntiles, s); break;
    case PSS_VARIATION_STRAND: e = launch_tile_one<KIND_VAR, PSS_VARIATION_STRAND>(a, ntiles, s); break;
    case PSS_MAPQ: e = launch_tilem_one<PSS_MAPQ>(a, ntiles, s); break;
    case PSS_MAPQ_STRAND: e = launch_tilem_one<PSS_MAPQ_STRAND>(a, ntiles, s); break;
    case PSS_BASEQ: e = a.lowq ? launch_tileq_one<KIND_BQ, PSS_BASEQ, true>(a, ntiles, s) : launch_tileq_one<KIND_BQ, PSS_BASEQ>(a, ntiles, s); break;
    case PSS_BASEQ_STRAND: e = a.lowq ? launch_tileq_one<KIND_BQ, PSS_BASEQ_STRAND, true>(a, ntiles, s)
                                      : launch_tileq_one<KIND_BQ, PSS_BASEQ_STRAND>(a, ntiles, s); break;
    case PSS_BASEQ_EXT: e = launch_tileq_one<KIND_BQX, PSS_BASEQ_EXT>(a, ntiles, s); break;
    case PSS_BASEQ_EXT_STRAND: e = launch_tileq_one<KIND_BQX, PSS_BASEQ_EXT_STRAND>(a, ntiles, s); break;
    case PSS_TLEN: e = launch_tilet_one<PSS_TLEN>(a, p.tlen, p.redo, ntiles, s); break;
    case PSS_TLEN_STRAND: e = launch_tilet_one<PSS_TLEN_STRAND>(a, p.tlen, p.redo, ntiles, s); break;
    default: e = cudaErrorInvalidValue;
    }
    if (e1) cudaEventRecord(e1, s);
#ifdef TILE_TIMING
    if (var && getenv("PSS_TILE_TIMING")) {
        std::vector<unsigned long long> h(need);
        cudaStreamSynchronize(s);
        cudaMemcpy(h.data(), tbuf, need * 8, cudaMemcpyDeviceToHost);
        double acc[8] = {0}; size_t cnt = 0;
        for (size_t t = 0; t + 8 <= need; t += 8) {
            if (!h[t]) continue;
            for (int k = 1; k < 8; ++k) acc[k] += (double)(h[t + k] - h[t + k - 1]);
            acc[0] += (double)(h[t + 7] - h[t]);
            ++cnt;
        }
        if (cnt) fprintf(stderr, "[tile timing, ns, %zu tiles] total %.0f | desc+issue %.0f | zero+sync %.0f | tma wait %.0f | reads %.0f | barrier %.0f | scan %.0f | finalise %.0f\n",
                         cnt, acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[4] / cnt, acc[5] / cnt, acc[6] / cnt, acc[7] / cnt);
    }
#endif
    return e;
}
