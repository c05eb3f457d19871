// The binned statistics: a read-parallel segmented reduction by read-start window that serves ANY SET of the
// five statistics from ONE pass over the read records.  sm_100a only.
//
// Reference: iter_binned_chrom (pysamstats/opt.pyx:2026-2102) rolls the window while pos > bin_end (strict), so
// window k >= 1 owns the reads with start + k*W < pos <= start + (k+1)*W, window 0 every read with
// pos <= start + W and the last window everything beyond.  Filters (min_mapq, no_dup opt.pyx:2066-2073) and the
// unmapped skip of every BinnedStat.recv (opt.pyx:1574, :1629, :1690, :1733, :1801) apply per read; the five
// recv / rec pairs are opt.pyx:1551-1819.
//
//   k_binned_acc<MASK>  the read stream is what bounds it, so it is read once whatever the statistics asked for, with
//                       16-byte vector loads: every thread walks EIGHT consecutive reads with the window's accumulators
//                       in registers (start sorted: a run rarely leaves its window), then the warp combines the runs
//                       window by window with full-mask REDUX instructions and lane j carries sum j to the window's
//                       accumulators (32-bit counts, 64-bit sums) -- see the kernel.  The window index is an exact
//                       multiply-high by ceil(2^64 / W), computed once per run, not a division per read.
//   k_binned_gc         one warp per window: G+C count of the window's reference bytes (coverage, coverage_ext)
//   k_binned_fin<MASK>  one thread per window: rec() of every statistic -- means / root mean squares, int32 columns
#include <limits.h>
#include <string.h>

#include "pss_internal.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ int dev_rootmean(unsigned long long sq, long long n) {
    return n > 0 ? (int)rint(sqrt((double)sq / (double)n)) : 0;   // opt.pyx:2159-2163
}
__device__ __forceinline__ int dev_mean(long long s, long long n) {
    return n > 0 ? (int)rint((double)s / (double)n) : 0;          // opt.pyx:2166-2170
}

// ---- accumulator layout of a statistic set (MASK bit i = enum pss_binned_stat i) --------------------------
// 32-bit count slots: n, n_pp | mapq0 | mate_unmapped, other_chr, same_strand, faceaway, softclipped, duplicate |
//                     the nine CIGAR op sums (C ints in the reference: they wrap)
// 64-bit sum slots:   sum mapq^2 | sum tlen, sum tlen (pp), sum tlen^2, sum tlen^2 (pp)
#define BM_COV (1 << PSS_COVERAGE_BINNED)
#define BM_EXT (1 << PSS_COVERAGE_EXT_BINNED)
#define BM_MQ (1 << PSS_MAPQ_BINNED)
#define BM_AL (1 << PSS_ALIGNMENT_BINNED)
#define BM_TL (1 << PSS_TLEN_BINNED)

__host__ __device__ constexpr int bslot_mq0(int) { return 2; }
__host__ __device__ constexpr int bslot_ext(int m) { return 2 + ((m & BM_MQ) ? 1 : 0); }
__host__ __device__ constexpr int bslot_al(int m) { return bslot_ext(m) + ((m & BM_EXT) ? 6 : 0); }
__host__ __device__ constexpr int binned_ncnt(int m) { return bslot_al(m) + ((m & BM_AL) ? 9 : 0); }
__host__ __device__ constexpr int bsum_tl(int m) { return (m & BM_MQ) ? 1 : 0; }
__host__ __device__ constexpr int binned_nsum(int m) { return bsum_tl(m) + ((m & BM_TL) ? 4 : 0); }

size_t binned_work_bytes(int mask, int64_t n_bins) {
    if (n_bins < 0) n_bins = 0;
    return (size_t)n_bins * (8 * (size_t)binned_nsum(mask) + 4 * (size_t)binned_ncnt(mask)) + 16;
}

struct BinnedArgs {
    long long start;
    long long n_bins;
    int window_size, min_mapq, no_dup, pad_;
    unsigned long long magic;      // ceil(2^64 / W): floor(x / W) == umul64hi(x, magic) for every x < 2^32
    unsigned long long *sums;      // [binned_nsum][n_bins]
    uint32_t *cnts;                // [binned_ncnt][n_bins]
    int32_t *out[PSS_N_BINNED_STATS];
    const uint8_t *ref;
    long long ref_start, ref_len;
};

// sum of v over the lanes of `mask` (all of them call), exact modulo 2^64
__device__ __forceinline__ unsigned long long seg_sum64(unsigned mask, unsigned long long v) {
    const unsigned long long m27 = (1ull << 27) - 1;
    unsigned long long s0 = __reduce_add_sync(mask, (unsigned)(v & m27));
    unsigned long long s1 = __reduce_add_sync(mask, (unsigned)((v >> 27) & m27));
    unsigned long long s2 = __reduce_add_sync(mask, (unsigned)(v >> 54));
    return s0 + (s1 << 27) + (s2 << 54);
}

// consecutive reads per thread: eight, or one when a statistic of the set walks every read's CIGAR (coverage_ext,
// alignment) -- neighbouring lanes then read neighbouring CIGAR words (measured: 0.24 vs 0.40 ms per 10 M reads)
__host__ __device__ constexpr int binned_r(int mask) { return (mask & (BM_AL | BM_EXT)) ? 1 : 8; }

// floor(rel / W) clamped to [0, nb - 1] (rel = pos - start - 1: the strict roll-over of opt.pyx:2050)
__device__ __forceinline__ int binned_window(long long rel, const BinnedArgs &A, unsigned magic_lo, unsigned magic_hi) {
    long long wl = rel < 0 ? 0 : ((rel >> 32) || !A.magic) ? rel / A.window_size
                   : (long long)(((unsigned long long)(unsigned)rel * magic_hi + __umulhi((unsigned)rel, magic_lo)) >> 32);
    if (wl > A.n_bins - 1) wl = A.n_bins - 1;
    return (int)wl;
}

// Thread g owns the BINNED_R consecutive reads [8 g - skew, 8 g - skew + 8) (skew = element misalignment of the
// record arrays, so the interior threads load their fields with 16-byte vector loads) and walks them in order with
// the window's accumulators in REGISTERS: the reads are start sorted, so a thread's run stays in one window seven
// times out of eight at 30x / 150 bp / W = 300; a window change inside the run adds the finished window's
// accumulators to memory directly.  What every thread holds at the end of its run (one window) is combined over the
// warp: one round per distinct window among the 32 lanes (256 reads: about five), a round being full-mask REDUX
// instructions over the lanes' contributions (zero outside the window) with lane j carrying the j-th sum to the
// window's accumulators -- ONE RED.32 and ONE RED.64 instruction per round, whatever the number of sums.
template <int MASK>
__global__ void __launch_bounds__(256)
k_binned_acc(pss_reads R, BinnedArgs A, long long n_groups, int skew, int vec_ok) {
    constexpr bool EXT = (MASK & BM_EXT) != 0, MQ = (MASK & BM_MQ) != 0, AL = (MASK & BM_AL) != 0, TL = (MASK & BM_TL) != 0;
    constexpr bool NEED_T = EXT || TL;
    constexpr int BINNED_R = binned_r(MASK);
    const int lane = threadIdx.x & 31;
    const long long nb = A.n_bins;
    // lane j carries the j-th count / sum of a round to its accumulator array (static assignment):
    //   counts: lanes 0-2 n, n_pp, mapq0 | 3-8 the six mate / clip classes | 9-17 CIGAR ops
    //   sums:   lane 0 mapq^2 | lanes 1-4 tlen sum, sum_pp, sum^2, sum^2_pp
    int cslot = -1, sslot = -1;
    if (lane < 2 || (MQ && lane == 2)) cslot = lane;
    if (EXT && lane >= 3 && lane < 9) cslot = bslot_ext(MASK) + lane - 3;
    if (AL && lane >= 9 && lane < 18) cslot = bslot_al(MASK) + lane - 9;
    if (MQ && lane == 0) sslot = 0;
    if (TL && lane >= 1 && lane < 5) sslot = bsum_tl(MASK) + lane - 1;
    uint32_t *const cacc = A.cnts + (long long)(cslot < 0 ? 0 : cslot) * nb;
    unsigned long long *const sacc = A.sums + (long long)(sslot < 0 ? 0 : sslot) * nb;
    const unsigned magic_lo = (unsigned)A.magic, magic_hi = (unsigned)(A.magic >> 32);
    const long long stride = (long long)gridDim.x * 256;
    for (long long gb = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32; gb < n_groups; gb += stride) {
        const long long i0 = (gb + lane) * BINNED_R - skew;
        int pos_[8], tlen_[8], mtid_[8];            // (BINNED_R of them used)
        unsigned flag_[8], mapq_[8];
        unsigned valid = 0;
        if (BINNED_R == 8 && vec_ok && i0 >= 0 && i0 + BINNED_R <= R.n_reads) {
            valid = (1u << BINNED_R) - 1u;
            const int4 p0 = __ldg((const int4 *)(R.pos + i0)), p1 = __ldg((const int4 *)(R.pos + i0) + 1);
            const uint4 f = __ldg((const uint4 *)(R.flag + i0));
            const uint2 m = __ldg((const uint2 *)(R.mapq + i0));
            pos_[0] = p0.x; pos_[1] = p0.y; pos_[2] = p0.z; pos_[3] = p0.w; pos_[4] = p1.x; pos_[5] = p1.y; pos_[6] = p1.z; pos_[7] = p1.w;
            flag_[0] = f.x & 0xffffu; flag_[1] = f.x >> 16; flag_[2] = f.y & 0xffffu; flag_[3] = f.y >> 16;
            flag_[4] = f.z & 0xffffu; flag_[5] = f.z >> 16; flag_[6] = f.w & 0xffffu; flag_[7] = f.w >> 16;
#pragma unroll
            for (int k = 0; k < 4; ++k) { mapq_[k] = (m.x >> (8 * k)) & 255u; mapq_[4 + k] = (m.y >> (8 * k)) & 255u; }
            if (NEED_T) {
                const int4 t0 = __ldg((const int4 *)(R.tlen + i0)), t1 = __ldg((const int4 *)(R.tlen + i0) + 1);
                tlen_[0] = t0.x; tlen_[1] = t0.y; tlen_[2] = t0.z; tlen_[3] = t0.w; tlen_[4] = t1.x; tlen_[5] = t1.y; tlen_[6] = t1.z; tlen_[7] = t1.w;
            }
            if (EXT) {
                const int4 t0 = __ldg((const int4 *)(R.mtid + i0)), t1 = __ldg((const int4 *)(R.mtid + i0) + 1);
                mtid_[0] = t0.x; mtid_[1] = t0.y; mtid_[2] = t0.z; mtid_[3] = t0.w; mtid_[4] = t1.x; mtid_[5] = t1.y; mtid_[6] = t1.z; mtid_[7] = t1.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < BINNED_R; ++k) {
                const long long i = i0 + k;
                const bool live = i >= 0 && i < R.n_reads;
                valid |= live ? 1u << k : 0u;
                pos_[k] = live ? __ldg(R.pos + i) : 0;
                flag_[k] = live ? (unsigned)__ldg(R.flag + i) : 4u;
                mapq_[k] = live ? (unsigned)__ldg(R.mapq + i) : 0u;
                tlen_[k] = (NEED_T && live) ? __ldg(R.tlen + i) : 0;
                mtid_[k] = (EXT && live) ? __ldg(R.mtid + i) : 0;
            }
        }
        // ---- the thread's run, accumulators in registers -------------------------------------------------
        int w = -1, wlo = 0, wend = INT_MIN;             // current window; its reads have wlo < pos <= wend (any read order works)
        unsigned n = 0, npp = 0, mq0 = 0, sq = 0;
        unsigned ext[6], al[9];
        long long ts = 0, tspp = 0;
        unsigned long long tq = 0, tqpp = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) ext[k] = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) al[k] = 0;
#pragma unroll
        for (int k = 0; k < BINNED_R; ++k) {
            if (!(valid >> k & 1)) continue;
            const int p = pos_[k];
            if (w < 0 || p > wend || p <= wlo) {
                if (w >= 0) {
                    // the run leaves window w: its accumulators go to memory directly
                    uint32_t *c = A.cnts + w;
                    unsigned long long *q = A.sums + w;
                    if (n) atomicAdd(c, n);
                    if (npp) atomicAdd(c + nb, npp);
                    if (MQ) { if (mq0) atomicAdd(c + (long long)bslot_mq0(MASK) * nb, mq0); if (sq) atomicAdd(q, (unsigned long long)sq); }
                    if (EXT) {
#pragma unroll
                        for (int j = 0; j < 6; ++j) if (ext[j]) atomicAdd(c + (long long)(bslot_ext(MASK) + j) * nb, ext[j]);
                    }
                    if (AL) {
#pragma unroll
                        for (int j = 0; j < 9; ++j) if (al[j]) atomicAdd(c + (long long)(bslot_al(MASK) + j) * nb, al[j]);
                    }
                    if (TL) {
                        if (ts) atomicAdd(q + (long long)(bsum_tl(MASK) + 0) * nb, (unsigned long long)ts);
                        if (tspp) atomicAdd(q + (long long)(bsum_tl(MASK) + 1) * nb, (unsigned long long)tspp);
                        if (tq) atomicAdd(q + (long long)(bsum_tl(MASK) + 2) * nb, tq);
                        if (tqpp) atomicAdd(q + (long long)(bsum_tl(MASK) + 3) * nb, tqpp);
                    }
                    n = npp = mq0 = sq = 0; ts = tspp = 0; tq = tqpp = 0;
#pragma unroll
                    for (int j = 0; j < 6; ++j) ext[j] = 0;
#pragma unroll
                    for (int j = 0; j < 9; ++j) al[j] = 0;
                }
                w = binned_window((long long)p - A.start - 1, A, magic_lo, magic_hi);
                const long long we = A.start + ((long long)w + 1) * A.window_size;
                wend = (w >= nb - 1 || we > INT_MAX) ? INT_MAX : (int)we;
                wlo = w == 0 ? INT_MIN : (int)(we - A.window_size);       // (< p, so it fits)
            }
            const unsigned flag = flag_[k];
            const int mapq = (int)mapq_[k];
            bool keep = !(flag & 4);                                   // every BinnedStat.recv
            if (A.min_mapq > 0 && mapq < A.min_mapq) keep = false;
            if (A.no_dup && (flag & 1024)) keep = false;
            if (!keep) continue;
            const bool rev = flag & 16, pp = flag & 2;
            n += 1; npp += pp ? 1u : 0u;                               // opt.pyx:1574-1584
            if (MQ) { mq0 += (mapq == 0); sq += (unsigned)(mapq * mapq); }   // opt.pyx:1690-1703 (8 x 255^2 fits)
            if (EXT) {                                                 // opt.pyx:1629-1662
                const bool mun = flag & 8, mrev = flag & 32;
                const int tl = tlen_[k];
                if (mun) ext[0]++;
                else if (R.tid != mtid_[k]) ext[1]++;
                else if (rev == mrev) ext[2]++;
                else if ((rev && tl > 0) || (!rev && tl < 0)) ext[3]++;
                int soft = 0;
                const long long i = i0 + k;
                for (int64_t c = R.cigar_off[i]; c < R.cigar_off[i + 1]; ++c) soft |= ((R.cigar[c] & 15) == OP_S);
                ext[4] += soft;
                ext[5] += (flag & 1024) ? 1u : 0u;
            }
            if (AL) {                                                  // opt.pyx:1733-1764: C int sums, they wrap
                const long long i = i0 + k;
                for (int64_t c = R.cigar_off[i]; c < R.cigar_off[i + 1]; ++c) {
                    const uint32_t cw = R.cigar[c];
                    const int op = cw & 15;
#pragma unroll
                    for (int o = 0; o <= OP_X; ++o) if (op == o) al[o] += (cw >> 4);
                }
            }
            if (TL) {                                                  // opt.pyx:1801-1819 (64-bit sums wrap like C)
                const long long t = tlen_[k];
                ts += t; tq += (unsigned long long)(t * t);
                if (pp) { tspp += t; tqpp += (unsigned long long)(t * t); }
            }
        }
        // ---- the runs' last windows, combined over the warp ------------------------------------------------
        unsigned remaining = __ballot_sync(FULL, w >= 0);
        while (remaining) {
            const int wl = __shfl_sync(FULL, w, __ffs(remaining) - 1);
            const bool in = w == wl;
            remaining &= ~__ballot_sync(FULL, in);
            unsigned cv = 0;                  // this lane's 32-bit count ...
            unsigned long long sv = 0;        // ... and 64-bit sum of the round
            {
                // n, n_pp, mapq0 <= 8 per lane: 10-bit fields hold the 32-lane sums
                const unsigned r = __reduce_add_sync(FULL, in ? (n | npp << 10 | mq0 << 20) : 0u);
                cv = (r >> (10 * (lane & 3))) & 1023u;              // lanes 0-2
            }
            if (MQ) sv = (unsigned long long)__reduce_add_sync(FULL, in ? sq : 0u);      // lane 0 (32 x 8 x 255^2 < 2^32)
            if (EXT) {
                const unsigned r2 = __reduce_add_sync(FULL, in ? (ext[0] | ext[1] << 10 | ext[2] << 20) : 0u);
                const unsigned r3 = __reduce_add_sync(FULL, in ? (ext[3] | ext[4] << 10 | ext[5] << 20) : 0u);
                if (lane >= 3) cv = lane < 6 ? (r2 >> (10 * (lane - 3))) & 1023u : (r3 >> (10 * ((lane - 6) & 3))) & 1023u;
            }
            if (AL) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const unsigned rk = __reduce_add_sync(FULL, in ? al[k] : 0u);
                    if (lane == 9 + k) cv = rk;
                }
            }
            if (TL) {
                const long long a = in ? ts : 0, b = in ? tspp : 0;
                const unsigned long long c = in ? tq : 0ull, d = in ? tqpp : 0ull;
                unsigned long long v1, v2, v3, v4;
                if (__all_sync(FULL, (c >> 32) == 0)) {
                    // ordinary template lengths: a lane's sum of squares below 2^32 bounds its |sum| by 2^17.5, so the
                    // 32-lane sums fit one REDUX (sign extended), the squares two 16-bit halves
                    v1 = (unsigned long long)(long long)(int)__reduce_add_sync(FULL, (unsigned)(int)a);
                    v2 = (unsigned long long)(long long)(int)__reduce_add_sync(FULL, (unsigned)(int)b);
                    v3 = (unsigned long long)__reduce_add_sync(FULL, (unsigned)c & 0xffffu) +
                         ((unsigned long long)__reduce_add_sync(FULL, (unsigned)c >> 16) << 16);
                    v4 = (unsigned long long)__reduce_add_sync(FULL, (unsigned)d & 0xffffu) +
                         ((unsigned long long)__reduce_add_sync(FULL, (unsigned)d >> 16) << 16);
                } else {
                    v1 = seg_sum64(FULL, (unsigned long long)a);
                    v2 = seg_sum64(FULL, (unsigned long long)b);
                    v3 = seg_sum64(FULL, c);
                    v4 = seg_sum64(FULL, d);
                }
                if (lane >= 1) sv = lane == 1 ? v1 : lane == 2 ? v2 : lane == 3 ? v3 : v4;
            }
            if (cslot >= 0 && cv) atomicAdd(cacc + wl, cv);
            if ((MQ || TL) && sslot >= 0 && sv) atomicAdd(sacc + wl, sv);
        }
    }
}

__global__ void __launch_bounds__(256)
k_binned_gc(BinnedArgs A, int32_t *__restrict__ out_a, int32_t *__restrict__ out_b) {
    const int lane = threadIdx.x & 31;
    const int64_t bin = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (bin >= A.n_bins) return;
    // rec(): gc_content(fafile.fetch(chrom, bin_start, bin_end)) opt.pyx:1561-1562
    const long long bstart = A.start + bin * (long long)A.window_size, bend = bstart + A.window_size;
    long long ws = bstart < 0 ? 0 : bstart, we = bend;
    const long long avail_hi = A.ref_start + A.ref_len;
    if (we > avail_hi) we = avail_hi;
    int c = 0;
    for (long long k = ws + lane; k < we; k += 32) {
        uint8_t ch = A.ref[k - A.ref_start] | 0x20;
        c += (ch == 'g' || ch == 'c');
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    // an empty window (beyond the FASTA) makes the reference's gc_content divide by zero
    // (opt.pyx:1538, ZeroDivisionError); reported as -1, the host layer raises
    if (lane == 0) {
        const int g = (we > ws) ? (int)rint((double)c * 100. / (double)(we - ws)) : -1;
        if (out_a) out_a[bin] = g;
        if (out_b) out_b[bin] = g;
    }
}

template <int MASK>
__global__ void __launch_bounds__(256)
k_binned_fin(BinnedArgs A) {
    const long long bin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= A.n_bins) return;
    const long long nb = A.n_bins;
    const uint32_t *c = A.cnts + bin;
    const unsigned long long *q = A.sums + bin;
    const int n = (int)c[0], npp = (int)c[nb];
    if (MASK & BM_COV) {                             // field 0 = gc (k_binned_gc)
        int32_t *o = A.out[PSS_COVERAGE_BINNED] + bin;
        o[nb] = n; o[2 * nb] = npp;
    }
    if (MASK & BM_EXT) {
        int32_t *o = A.out[PSS_COVERAGE_EXT_BINNED] + bin;
        o[nb] = n; o[2 * nb] = npp;
#pragma unroll
        for (int k = 0; k < 6; ++k) o[(3 + k) * nb] = (int)c[(long long)(bslot_ext(MASK) + k) * nb];
    }
    if (MASK & BM_MQ) {
        int32_t *o = A.out[PSS_MAPQ_BINNED] + bin;
        o[0] = n; o[nb] = (int)c[(long long)bslot_mq0(MASK) * nb];
        o[2 * nb] = dev_rootmean(q[0], n);
    }
    if (MASK & BM_AL) {
        // reads_all bases_all M I D N S H P = X (config.py dtype_alignment_binned); int accumulators in the reference
        int32_t *o = A.out[PSS_ALIGNMENT_BINNED] + bin;
        uint32_t ops[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) ops[k] = c[(long long)(bslot_al(MASK) + k) * nb];
        o[0] = n;
        o[nb] = (int32_t)(ops[OP_M] + ops[OP_I] + ops[OP_S] + ops[OP_EQ] + ops[OP_X]);
#pragma unroll
        for (int k = 0; k < 9; ++k) o[(2 + k) * nb] = (int32_t)ops[k];
    }
    if (MASK & BM_TL) {
        int32_t *o = A.out[PSS_TLEN_BINNED] + bin;
        const long long s = (long long)q[(long long)(bsum_tl(MASK) + 0) * nb], spp = (long long)q[(long long)(bsum_tl(MASK) + 1) * nb];
        const unsigned long long s2 = q[(long long)(bsum_tl(MASK) + 2) * nb], s2pp = q[(long long)(bsum_tl(MASK) + 3) * nb];
        o[0] = n; o[nb] = npp;
        o[2 * nb] = dev_mean(s, n); o[3 * nb] = dev_mean(spp, npp);
        o[4 * nb] = dev_rootmean(s2, n); o[5 * nb] = dev_rootmean(s2pp, npp);
    }
}

// ------------------------------------------------------------------------ launcher

#define LAUNCHED() do { if (launches) ++*launches; } while (0)

template <int MASK>
static void launch_binned_mask(const pss_reads &R, const BinnedArgs &A, int sm_count, cudaStream_t s, int64_t *launches,
                               const BinnedEvents *ev) {
    if (R.n_reads > 0) {
        // thread g owns the reads [8 g - skew, 8 g - skew + 8): with the common element misalignment of the record
        // arrays as skew, the interior threads' 16-byte vector loads are aligned
        const int skew = (int)(((uintptr_t)R.pos >> 2) & 7);     // (used by the eight-reads-per-thread instantiations only)
        const bool need_t = (MASK & (BM_EXT | BM_TL)) != 0, ext = (MASK & BM_EXT) != 0;
        const int vec_ok = ((uintptr_t)R.pos & 3) == 0 && (int)(((uintptr_t)R.flag >> 1) & 7) == skew && ((uintptr_t)R.flag & 1) == 0 &&
                           (int)((uintptr_t)R.mapq & 7) == skew &&
                           (!need_t || ((int)(((uintptr_t)R.tlen >> 2) & 7) == skew && ((uintptr_t)R.tlen & 3) == 0)) &&
                           (!ext || ((int)(((uintptr_t)R.mtid >> 2) & 7) == skew && ((uintptr_t)R.mtid & 3) == 0));
        constexpr int BINNED_R = binned_r(MASK);
        const long long n_groups = BINNED_R == 8 ? (R.n_reads + skew + 7) / 8 : R.n_reads;
        const long long n_ctas = (n_groups + 255) / 256;
        const long long cap = (long long)sm_count * 8;
        if (ev && ev->acc0) cudaEventRecord(ev->acc0, s);
        k_binned_acc<MASK><<<(unsigned)(n_ctas < cap ? n_ctas : cap), 256, 0, s>>>(R, A, n_groups, (vec_ok && BINNED_R == 8) ? skew : 0,
                                                                                 vec_ok && BINNED_R == 8);
        if (ev && ev->acc1) cudaEventRecord(ev->acc1, s);
        LAUNCHED();
    }
    if (ev && ev->fin0) cudaEventRecord(ev->fin0, s);
    k_binned_fin<MASK><<<(unsigned)((A.n_bins + 255) / 256), 256, 0, s>>>(A);
    if (ev && ev->fin1) cudaEventRecord(ev->fin1, s);
    LAUNCHED();
}

// `work`: binned_work_bytes(mask, n_bins) bytes of device scratch (the windows' accumulators); outs[i] (device) for
// every statistic i of `mask`.  P carries the common request (stat is ignored).
cudaError_t launch_binned_multi(const pss_reads &R, const pss_binned_params &P, int mask, int32_t *const *outs, void *work,
                                cudaStream_t s, int64_t *launches, const BinnedEvents *ev) {
    if (P.n_bins <= 0 || mask <= 0 || mask >= (1 << PSS_N_BINNED_STATS)) return cudaSuccess;
    BinnedArgs A;
    memset(&A, 0, sizeof(A));
    A.start = P.start; A.n_bins = P.n_bins; A.window_size = P.window_size; A.min_mapq = P.min_mapq; A.no_dup = P.no_dup;
    // ceil(2^64 / W); W == 1 has no 64-bit magic (0 selects the division in the kernel)
    A.magic = P.window_size > 1 ? ~0ull / (unsigned long long)P.window_size + 1ull : 0ull;
    A.sums = (unsigned long long *)work;
    A.cnts = (uint32_t *)(A.sums + (size_t)binned_nsum(mask) * (size_t)P.n_bins);
    for (int i = 0; i < PSS_N_BINNED_STATS; ++i) A.out[i] = (mask >> i & 1) ? outs[i] : nullptr;
    A.ref = P.ref; A.ref_start = P.ref_start; A.ref_len = P.ref_len;
    cudaError_t e = cudaMemsetAsync(work, 0, binned_work_bytes(mask, P.n_bins), s);
    if (e != cudaSuccess) return e;
    int dev = 0, sm_count = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    switch (mask) {
#define CASE(m) case m: launch_binned_mask<m>(R, A, sm_count, s, launches, ev); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
        CASE(15) CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23) CASE(24) CASE(25) CASE(26) CASE(27)
        CASE(28) CASE(29) CASE(30) CASE(31)
#undef CASE
    default: break;
    }
    if (mask & (BM_COV | BM_EXT)) {
        if (ev && ev->gc0) cudaEventRecord(ev->gc0, s);
        k_binned_gc<<<(unsigned)((P.n_bins * 32 + 255) / 256), 256, 0, s>>>(A, A.out[PSS_COVERAGE_BINNED], A.out[PSS_COVERAGE_EXT_BINNED]);
        if (ev && ev->gc1) cudaEventRecord(ev->gc1, s);
        LAUNCHED();
    }
    return cudaGetLastError();
}
