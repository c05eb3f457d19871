// Per-read preparation, depth-cap admission, mate-overlap quality rewrite, tile ranges,
// GC profile and the binned segmented reduction.  sm_100a only.
//
// Reference citations are relative to the reference repository root; "A.x" refers to
// SURVEY.md Appendix A (the htslib semantics behind AlignmentFile.pileup, which the
// reference consumes at pysamstats/opt.pyx:1853-1854).
#include <limits.h>
#include <string.h>

#include "pss_internal.cuh"

#define FULL 0xffffffffu

// ------------------------------------------------------------------ block primitives

__device__ __forceinline__ int warp_sum(int v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// all-threads block reductions for blockDim.x <= 1024 (multiple of 32); s = 33 ints of smem
__device__ int block_sum(int v, int *s) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s[w] = v;
    __syncthreads();
    int t = (lane < nw) ? s[lane] : 0;
    return warp_sum(t);
}
__device__ int block_min(int v, int *s) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_min(v);
    __syncthreads();
    if (lane == 0) s[w] = v;
    __syncthreads();
    int t = (lane < nw) ? s[lane] : INT_MAX;
    return warp_min(t);
}
// exclusive prefix sum over the block in thread order; *total receives the block sum
__device__ int block_excl_scan(int v, int *s, int *total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s[w] = inc;
    __syncthreads();
    int ws = (lane < nw) ? s[lane] : 0;
    int wi = ws;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, wi, o);
        if (lane >= o) wi += t;
    }
    int wexcl = __shfl_sync(FULL, wi - ws, w);
    *total = __shfl_sync(FULL, wi, 31);
    return wexcl + inc - v;
}

// ------------------------------------------------------------------------- k_prep
//
// One thread per read: walks the CIGAR once (bam_endpos, is_softclipped opt.pyx:2133-2140),
// applies the stepper flag filter (A.0), folds the read-level pysamstats filters
// (min_mapq opt.pyx:1897-1899, no_dup :1907-1909) and the mate-class chain of
// CoverageExt.recv (opt.pyx:266-279) into one class word.

__global__ void __launch_bounds__(256)
k_prep(pss_reads R, int stepper, int min_mapq, int no_dup, RInfo *__restrict__ ri, SmallState *st) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int span = 0;
    if (i < R.n_reads) {
        int pos = R.pos[i];
        uint32_t flag = R.flag[i];
        int mapq = R.mapq[i];
        int tlen = R.tlen[i];
        int mtid = R.mtid[i];
        int lseq = R.l_seq[i];
        int64_t co = R.cigar_off[i];
        int nc = (int)(R.cigar_off[i + 1] - co);
        int64_t reflen = 0;
        int soft = 0, first_op = -1;
        for (int k = 0; k < nc; ++k) {
            uint32_t w = R.cigar[co + k];
            int op = w & 15;
            if (k == 0) first_op = op;
            if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) reflen += (w >> 4);
            if (op == OP_S) soft = 1;
        }
        bool unmapped = (flag & 4) != 0;
        bool noref = unmapped || reflen == 0;
        int end = pos + (noref ? 1 : (int)reflen);
        bool pushed;
        if (stepper == PSS_STEPPER_NOFILTER) pushed = true;
        else {
            pushed = !(flag & 0x704);
            if (stepper == PSS_STEPPER_SAMTOOLS && (flag & 1) && !(flag & 2)) pushed = false;
        }
        bool admitted = pushed && !unmapped && R.tid >= 0;
        bool rev = flag & 16, pp = flag & 2, mun = flag & 8, mrev = flag & 32, dup = flag & 1024;
        uint32_t cls = 0;
        if (rev) cls |= CLS_REV;
        if (pp) cls |= CLS_PP;
        if (admitted) {
            cls |= CLS_ADMITTED;
            if (!(min_mapq > 0 && mapq < min_mapq) && !(no_dup && dup)) cls |= CLS_COUNTED;
        } else if (pushed) cls |= CLS_PUSHDEL;
        uint32_t ext = 1u | (pp ? 2u : 0u);
        if (mun) ext |= 4u;
        else if (R.tid != mtid) ext |= 8u;
        else if (rev == mrev) ext |= 16u;
        else if ((rev && tlen > 0) || (!rev && tlen < 0)) ext |= 32u;
        if (soft) ext |= 64u;
        if (dup) ext |= 128u;
        cls |= ext << CLS_EXT_SHIFT;
        if (!mun && R.tid == mtid) cls |= CLS_PAIRED;
        if (nc == 1 && op_is_match(first_op)) cls |= CLS_SIMPLE;
        if (lseq > 0) cls |= CLS_HASSEQ;
        if (noref) cls |= CLS_NOREF;
        cls |= (uint32_t)(mapq & 255) << CLS_MAPQ_SHIFT;
        long long at = tlen < 0 ? -(long long)tlen : (long long)tlen;
        if (!mun && pp && at < 2ll * lseq) cls |= CLS_OVL;
        RInfo r;
        r.pos = pos; r.end = end; r.cls = cls; r.n_cigar = (uint32_t)nc;
        r.cigar_off = co; r.seq_off = R.seq_off[i];
        ri[i] = r;
        span = end - pos;
        if (i > 0 && R.pos[i - 1] > pos) atomicOr(&st->unsorted, 1);
    }
    // one atomic per warp at most, and only while the running maximum still grows (a
    // same-address atomic per warp would serialise in L2)
    span = warp_max(span);
    if ((threadIdx.x & 31) == 0 && span > *(volatile int32_t *)&st->maxspan) atomicMax(&st->maxspan, span);
}

// -------------------------------------------------------------- depth cap (A.2)
//
// bam_plp_push refuses read i iff it_pos == pos_i and the list already holds >= maxcnt
// nodes.  Closed form: within each same-start group of admitted-so-far reads (file order)
// the read of rank r is kept iff r == 0 or L0 + r < maxcnt, L0 = kept earlier reads with
// end >= pos.  A read is provably unconstrained ("safe") when the read maxcnt places
// earlier starts more than maxspan before it: then fewer than maxcnt earlier reads can
// still be live.  k_safe marks the index range that is not provably safe; k_admit_slow
// resolves that range exactly, group by group, with one CTA.

__global__ void __launch_bounds__(256)
k_safe(const int32_t *__restrict__ pos, int64_t n, int maxcnt, SmallState *st) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + maxcnt;
    int lo = INT_MAX, hi = -1;
    if (i < n) {
        if (pos[i] - pos[i - maxcnt] <= st->maxspan) { lo = (int)i; hi = (int)i; }
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0 && hi >= 0) {
        atomicMin(&st->first_unsafe, lo);
        atomicMax(&st->last_unsafe, hi);
    }
}

__global__ void __launch_bounds__(1024)
k_admit_slow(RInfo *ri, int64_t n, int maxcnt, const SmallState *st) {
    __shared__ int s_red[33];
    __shared__ long long s_i;
    int first = st->first_unsafe, last = st->last_unsafe;
    if (first > last) return;
    const int tid = threadIdx.x;
    const int maxspan = st->maxspan;
    volatile RInfo *v = ri;
    if (tid == 0) {
        long long i = first;
        int p = v[i].pos;
        while (i > 0 && v[i - 1].pos == p) --i;
        s_i = i;
    }
    __syncthreads();
    long long i = s_i;
    while (i <= last) {
        const int p = v[i].pos;
        // end of the same-start group
        long long gend;
        for (long long base = i;; base += 1024) {
            long long j = base + tid;
            int differs = (j >= n || v[j].pos != p) ? 1 : 0;
            if (__syncthreads_or(differs)) {
                gend = base + block_min(differs ? tid : 1024, s_red);
                break;
            }
        }
        // L0: kept earlier reads still live when the iterator sits at p (end >= p)
        int c = 0;
        for (long long hi = i;; hi -= 1024) {
            long long j = hi - 1 - tid;
            int stop = 0;
            if (j >= 0) {
                int jp = v[j].pos;
                if (jp < p - maxspan) stop = 1;
                else if ((v[j].cls & CLS_ADMITTED) && v[j].end >= p) c++;
            } else stop = 1;
            if (__syncthreads_or(stop)) break;
        }
        const int L0 = block_sum(c, s_red);
        // ranks within the group, file order
        int rank_base = 0;
        for (long long base = i; base < gend; base += 1024) {
            long long j = base + tid;
            uint32_t cls = 0;
            int elig = 0;
            if (j < gend) { cls = v[j].cls; elig = (cls & CLS_ADMITTED) ? 1 : 0; }
            int total;
            int r = rank_base + block_excl_scan(elig, s_red, &total);
            if (elig && r > 0 && (long long)L0 + r >= maxcnt)
                v[j].cls = (cls & ~(CLS_ADMITTED | CLS_COUNTED)) | CLS_DROPPED;
            rank_base += total;
        }
        __threadfence_block();
        __syncthreads();
        i = gend;
    }
}

__global__ void k_export_admitted(const RInfo *__restrict__ ri, int64_t n, uint8_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (ri[i].cls & CLS_ADMITTED) ? 1 : 0;
}

// -------------------------------------------- mate-overlap quality rewrite (A.3)
//
// htslib keeps a QNAME-keyed table while reads are pushed: a qualifying read either parks
// itself under its name or, if a read is parked there, is tweaked against it and the key
// is deleted; keys are also deleted when a same-name read is refused / unmapped, or when
// any same-name node expires from the list.  Because every event on a key involves only
// reads of that name, the table can be replayed independently per name: k_pair walks each
// name chain (name_prev / name_next links, file order) with the key state in a register.
// Expiry of node x happens before the push of read r iff x.end <= P_r - 1 where P_r is the
// start of the last accepted read before r (the iterator has emitted every column < P_r).

__global__ void k_name_next(const int32_t *__restrict__ prev, int64_t n, int32_t *next) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && prev[i] >= 0) next[prev[i]] = (int32_t)i;
}

__device__ __forceinline__ long long prev_accepted_pos(const RInfo *ri, int64_t r) {
    for (int64_t j = r - 1; j >= 0; --j)
        if (ri[j].cls & CLS_ADMITTED) return ri[j].pos;
    return LLONG_MIN / 4;
}

__global__ void __launch_bounds__(256)
k_pair(const RInfo *__restrict__ ri, const int32_t *__restrict__ prev, const int32_t *__restrict__ next,
       int64_t n, int32_t *pair_first) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || prev[i] >= 0) return;
    int stored = -1;
    long long P_prev = LLONG_MIN / 4;
    for (int r = (int)i; r >= 0; r = next[r]) {
        uint32_t cls = ri[r].cls;
        bool accepted = cls & CLS_ADMITTED;
        bool del_ev = cls & (CLS_PUSHDEL | CLS_DROPPED);
        if (!accepted && !del_ev) continue;  // filtered by the stepper: never reaches push
        long long P = prev_accepted_pos(ri, r);
        if (stored >= 0) {
            for (int x = (int)i; x != r; x = next[x]) {
                if (!(ri[x].cls & CLS_ADMITTED)) continue;
                long long e = ri[x].end;
                if (e > P_prev - 1 && e <= P - 1) { stored = -1; break; }
            }
        }
        P_prev = P;
        if (del_ev) { stored = -1; continue; }
        if (cls & CLS_OVL) {
            if (stored < 0) stored = r;
            else { pair_first[r] = stored; stored = -1; }
        }
    }
}

// the two cursor functions of htslib's tweak_overlap_quality, old-htslib flavour
// (in-block counter reset to 0, A.3 "quirk")
struct Cur { int ci, icig, iseq, iref; };

__device__ bool cur_set(const uint32_t *cg, int nc, int want, Cur &c) {
    c.ci = c.icig = c.iseq = c.iref = 0;
    if (want < 0) return false;
    while (c.ci < nc) {
        uint32_t w = cg[c.ci];
        int op = w & 15, len = (int)(w >> 4);
        if (op_is_match(op)) {
            want -= len;
            if (want < 0) { c.icig = len + want; c.iseq += c.icig; c.iref += c.icig; return true; }
            c.iseq += len; c.iref += len;
        } else if (op == OP_S || op == OP_I) c.iseq += len;
        else if (op == OP_D || op == OP_N) { want -= len; if (want < 0) want = 0; c.iref += len; }
        c.ci++; c.icig = 0;
    }
    return false;
}

__device__ bool cur_next(const uint32_t *cg, int nc, Cur &c) {
    while (c.ci < nc) {
        uint32_t w = cg[c.ci];
        int op = w & 15, len = (int)(w >> 4);
        if (op_is_match(op)) {
            if (c.icig >= len - 1) { c.icig = 0; c.ci++; continue; }
            c.iseq++; c.icig++; c.iref++;
            return true;
        }
        if (op == OP_D || op == OP_N) c.iref += len;
        else if (op == OP_I || op == OP_S) c.iseq += len;
        c.ci++; c.icig = 0;
    }
    c.iseq = c.iref = -1;
    return false;
}

__device__ __forceinline__ int seq_code(const uint8_t *s, int k) { return (s[k >> 1] >> ((~k & 1) << 2)) & 15; }

__global__ void __launch_bounds__(128)
k_tweak(const RInfo *__restrict__ ri, const uint32_t *__restrict__ cigar, const uint8_t *__restrict__ seq4,
        const int64_t *__restrict__ qual_off, const int32_t *__restrict__ pair_first, int64_t n, uint8_t *qual) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    int a = pair_first[b];
    if (a < 0) return;
    RInfo ra = ri[a], rb = ri[b];
    const uint32_t *ca = cigar + ra.cigar_off, *cb = cigar + rb.cigar_off;
    int na = (int)ra.n_cigar, nb = (int)rb.n_cigar;
    const uint8_t *sa = seq4 + ra.seq_off, *sb = seq4 + rb.seq_off;
    uint8_t *qa = qual + qual_off[a], *qb = qual + qual_off[b];
    int apos = ra.pos, bpos = rb.pos;
    int iref = bpos;
    Cur A, B;
    if (!cur_set(ca, na, iref - apos, A)) return;
    if (!cur_set(cb, nb, iref - bpos, B)) return;
    bool oka = true, okb = true;
    for (;;) {
        while (A.iref >= 0 && A.iref < iref - apos) oka = cur_next(ca, na, A);
        if (!oka) return;
        iref = max(iref, A.iref + apos);
        while (B.iref >= 0 && B.iref < iref - bpos) okb = cur_next(cb, nb, B);
        if (!okb) return;
        iref = max(iref, B.iref + bpos);
        iref++;
        if (A.iref + apos != B.iref + bpos) continue;
        int va = qa[A.iseq], vb = qb[B.iseq];
        if (seq_code(sa, A.iseq) == seq_code(sb, B.iseq)) {
            int s = va + vb;
            qa[A.iseq] = (uint8_t)(s > 200 ? 200 : s);
            qb[B.iseq] = 0;
        } else if (va >= vb) {
            qa[A.iseq] = (uint8_t)(va * 4 / 5);
            qb[B.iseq] = 0;
        } else {
            qb[B.iseq] = (uint8_t)(vb * 4 / 5);
            qa[A.iseq] = 0;
        }
    }
}

// -------------------------------------------------------------------- tile ranges
//
// Candidate reads of tile [t0, t1): reads are start-sorted, so they are the index range
// [first pos >= t0 - maxspan + 1, first pos >= t1).

__device__ __forceinline__ int64_t lower_bound_pos(const RInfo *ri, int64_t n, long long key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((long long)ri[mid].pos < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_tile_ranges(const RInfo *__restrict__ ri, int64_t n, const SmallState *st, int64_t start,
              int64_t ntiles, int tile, int2 *rng) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    long long t0 = start + t * tile, t1 = t0 + tile;
    int2 r;
    r.x = (int)lower_bound_pos(ri, n, t0 - st->maxspan + 1);
    r.y = (int)lower_bound_pos(ri, n, t1);
    rng[t] = r;
}

// ---------------------------------------------------------------------- GC profile
//
// coverage_gc's per-position window GC% (CoverageGC.rec opt.pyx:1498-1509 + gc_content
// :1527-1539): window [max(0, p - off), p - off + size) clipped to the contig, percent of
// g/c (any case) over the ACTUAL window length, round-half-even.  Each CTA builds a
// prefix sum of its reference span in shared memory.

#define GC_BLK 2048
#define GC_REGION 8192
#define GC_TAB 1024

// per-byte "is g / c (either case)" flags of four reference bytes, gathered into a nibble
__device__ __forceinline__ unsigned gc_nibble(unsigned w) {
    unsigned y = ((w | 0x20202020u) & 0xFBFBFBFBu) ^ 0x63636363u;        // 'g' = 0x67, 'c' = 0x63: zero byte <=> g / c
    unsigned z = ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;   // bit 7 of every zero byte
    return (((z >> 7) * 0x01020408u) >> 24);
}

// int(np.around(cnt * 100. / n)) in integers: the correctly rounded double quotient of two integers below 2^45
// cannot land on a half unless the exact quotient is one (its distance from a half is >= 1 / 2n), and rounding
// is monotone, so the double expression of gc_content and the exact rational agree.
__device__ __forceinline__ int gc_percent(unsigned long long cnt, unsigned long long n) {
    unsigned long long v = cnt * 100ull, q = v / n, r = v - q * n;
    return (int)(q + ((2 * r > n) || (2 * r == n && (q & 1))));
}
__device__ __forceinline__ int gc_percent32(unsigned cnt, unsigned n) {
    unsigned v = cnt * 100u, q = v / n, r = v - q * n;
    return (int)(q + ((2 * r > n) || (2 * r == n && (q & 1))));
}

// A CTA turns its reference span into one 32-bit flag mask + one running count per 32 bytes (16-byte vector
// loads); a window count is two of those lookups, the percentage a table lookup while the window is not
// clipped by the contig ends (else an integer division); four positions per thread and 128-bit stores.
__global__ void __launch_bounds__(256)
k_gc(const uint8_t *__restrict__ ref, int64_t ref_start, int64_t ref_len, int64_t contig_len,
     int64_t start, int64_t end, int wsize, int woff, int32_t *__restrict__ out) {
    __shared__ int s_pre[257];
    __shared__ unsigned s_mask[257];
    __shared__ int s_red[33];
    __shared__ uint8_t s_tab[GC_TAB + 1];
    const int64_t b0 = start + (int64_t)blockIdx.x * GC_BLK;
    int64_t b1 = b0 + GC_BLK;
    if (b1 > end) b1 = end;
    int64_t rlo = b0 - woff, rhi = (b1 - 1) - woff + wsize;  // region of reference needed
    const bool unclipped = rlo >= 0 && rhi <= contig_len && wsize > 0 && wsize <= GC_TAB;
    if (rlo < 0) rlo = 0;
    if (rhi > contig_len) rhi = contig_len;
    const int64_t rn = rhi - rlo;
    // smem index i <-> address abase + i; the region starts at index shift
    const long long raddr = (long long)(uintptr_t)ref, gaddr = raddr + (rlo - ref_start);
    const long long abase = gaddr & ~15ll;
    const int shift = (int)(gaddr - abase);
    const bool fast = rn > 0 && shift + rn <= GC_REGION;
    if (fast) {
        const long long a = abase + 32ll * threadIdx.x;
        unsigned m = 0;
        if (a < gaddr + rn && a + 32 > gaddr) {
            if (a >= raddr && a + 32 <= raddr + ref_len) {
                const uint4 *vp = (const uint4 *)(uintptr_t)a;
                uint4 v0 = __ldg(vp), v1 = __ldg(vp + 1);
                m = gc_nibble(v0.x) | gc_nibble(v0.y) << 4 | gc_nibble(v0.z) << 8 | gc_nibble(v0.w) << 12 |
                    gc_nibble(v1.x) << 16 | gc_nibble(v1.y) << 20 | gc_nibble(v1.z) << 24 | gc_nibble(v1.w) << 28;
            } else {
                for (int j = 0; j < 32; ++j) {
                    long long ad = a + j;
                    if (ad >= raddr && ad < raddr + ref_len) {
                        uint8_t ch = *(const uint8_t *)(uintptr_t)ad | 0x20;
                        m |= (unsigned)(ch == 'g' || ch == 'c') << j;
                    }
                }
            }
        }
        int total;
        int run = block_excl_scan(__popc(m), s_red, &total);
        s_pre[threadIdx.x] = run;
        s_mask[threadIdx.x] = m;
        if (threadIdx.x == 0) { s_pre[256] = total; s_mask[256] = 0; }
        if (unclipped)
            for (int c = threadIdx.x; c <= wsize; c += 256) s_tab[c] = (uint8_t)gc_percent32((unsigned)c, (unsigned)wsize);
        __syncthreads();
    }
    auto prefix = [&](int i) { return s_pre[i >> 5] + __popc(s_mask[i >> 5] & ((1u << (i & 31)) - 1u)); };
    auto gc_at = [&](int64_t p) -> int {
        int64_t ws = p - woff, we = ws + wsize;
        if (ws < 0) ws = 0;
        if (we > contig_len) we = contig_len;
        if (we <= ws) return -1;
        if (fast) {
            int cnt = prefix((int)(we - rlo) + shift) - prefix((int)(ws - rlo) + shift);
            return gc_percent32((unsigned)cnt, (unsigned)(we - ws));
        }
        unsigned long long cnt = 0;
        for (int64_t k = ws; k < we; ++k) {
            int64_t ro = k - ref_start;
            uint8_t ch = (ro >= 0 && ro < ref_len) ? (ref[ro] | 0x20) : 0;
            cnt += (ch == 'g' || ch == 'c');
        }
        return gc_percent(cnt, (unsigned long long)(we - ws));
    };
    const int nb = (int)(b1 - b0);
    int32_t *o = out + (b0 - start);
    const bool vec = (((uintptr_t)o) & 15) == 0;
    int q0 = 0;
    if (fast && unclipped && vec) {
        // eight positions per thread: one window count from the running counts, then the window slides over flag bits
        // held in two registers (the 32 bases from the window's first / one-past-last base on)
        const int n8 = nb & ~7;
        for (int q = threadIdx.x * 8; q < n8; q += 2048) {
            const int lo = shift + q, hi = lo + wsize;
            const unsigned wl = __funnelshift_r(s_mask[lo >> 5], s_mask[(lo >> 5) + 1], lo & 31);
            const unsigned wh = __funnelshift_r(s_mask[hi >> 5], s_mask[(hi >> 5) + 1], hi & 31);
            int c = prefix(hi) - prefix(lo);
            int g[8];
            g[0] = s_tab[c];
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                c += (int)((wh >> (k - 1)) & 1u) - (int)((wl >> (k - 1)) & 1u);
                g[k] = s_tab[c];
            }
            __stcs((int4 *)(o + q), make_int4(g[0], g[1], g[2], g[3]));
            __stcs((int4 *)(o + q + 4), make_int4(g[4], g[5], g[6], g[7]));
        }
        q0 = n8;
    }
    for (int q = q0 + threadIdx.x * 4; q < nb; q += 1024) {
        if (vec && q + 4 <= nb) {
            int4 g;
            if (fast && unclipped) {
                const int lo = (int)(b0 - woff - rlo) + shift + q;     // window start of position b0 + q
                int c0 = prefix(lo + wsize) - prefix(lo);
                // the window slides by one base per position
                auto bit = [&](int i) { return (int)((s_mask[i >> 5] >> (i & 31)) & 1u); };
                int c1 = c0 - bit(lo) + bit(lo + wsize);
                int c2 = c1 - bit(lo + 1) + bit(lo + 1 + wsize);
                int c3 = c2 - bit(lo + 2) + bit(lo + 2 + wsize);
                g = make_int4(s_tab[c0], s_tab[c1], s_tab[c2], s_tab[c3]);
            } else {
                g = make_int4(gc_at(b0 + q), gc_at(b0 + q + 1), gc_at(b0 + q + 2), gc_at(b0 + q + 3));
            }
            __stcs((int4 *)(o + q), g);
        } else {
            for (int j = q; j < q + 4 && j < nb; ++j) o[j] = gc_at(b0 + j);
        }
    }
}

// ------------------------------------------------------------------------ launchers

#define LAUNCHED() do { if (launches) ++*launches; } while (0)

cudaError_t launch_prep(const pss_reads &R, int stepper, int min_mapq, int no_dup, RInfo *ri,
                        SmallState *st, cudaStream_t s, int64_t *launches) {
    SmallState init;
    memset(&init, 0, sizeof(init));
    init.maxspan = 1; init.unsorted = 0; init.first_unsafe = INT_MAX; init.last_unsafe = -1;
    cudaError_t e = cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    if (R.n_reads > 0) {
        unsigned grid = (unsigned)((R.n_reads + 255) / 256);
        k_prep<<<grid, 256, 0, s>>>(R, stepper, min_mapq, no_dup, ri, st);
        LAUNCHED();
    }
    return cudaGetLastError();
}

cudaError_t launch_admit(RInfo *ri, const int32_t *pos, int64_t n, int maxcnt, SmallState *st, cudaStream_t s,
                         int64_t *launches) {
    if (n > maxcnt) {
        unsigned grid = (unsigned)((n - maxcnt + 255) / 256);
        k_safe<<<grid, 256, 0, s>>>(pos, n, maxcnt, st);
        LAUNCHED();
        k_admit_slow<<<1, 1024, 0, s>>>(ri, n, maxcnt, st);   // returns at once when nothing is unsafe
        LAUNCHED();
    }
    return cudaGetLastError();
}

cudaError_t launch_export_admitted(const RInfo *ri, int64_t n, uint8_t *admitted, cudaStream_t s,
                                   int64_t *launches) {
    if (n > 0) {
        k_export_admitted<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ri, n, admitted);
        LAUNCHED();
    }
    return cudaGetLastError();
}

cudaError_t launch_overlap(const pss_reads &R, const RInfo *ri, uint8_t *qual_work, int32_t *pair_first,
                           int32_t *name_next, cudaStream_t s, int64_t *launches) {
    int64_t n = R.n_reads;
    if (n == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaMemsetAsync(pair_first, 0xff, sizeof(int32_t) * n, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(name_next, 0xff, sizeof(int32_t) * n, s)) != cudaSuccess) return e;
    unsigned grid = (unsigned)((n + 255) / 256);
    k_name_next<<<grid, 256, 0, s>>>(R.name_prev, n, name_next);
    LAUNCHED();
    k_pair<<<grid, 256, 0, s>>>(ri, R.name_prev, name_next, n, pair_first);
    LAUNCHED();
    if (qual_work) {     // NULL: the pairs only (pss_pairs_host)
        k_tweak<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ri, R.cigar, R.seq4, R.qual_off, pair_first, n, qual_work);
        LAUNCHED();
    }
    return cudaGetLastError();
}

cudaError_t launch_tile_ranges(const RInfo *ri, int64_t n, const SmallState *st, int64_t start,
                               int64_t ntiles, int tile, int2 *rng, cudaStream_t s, int64_t *launches) {
    if (ntiles > 0) {
        k_tile_ranges<<<(unsigned)((ntiles + 255) / 256), 256, 0, s>>>(ri, n, st, start, ntiles, tile, rng);
        LAUNCHED();
    }
    return cudaGetLastError();
}

cudaError_t launch_gc(const uint8_t *ref, int64_t ref_start, int64_t ref_len, int64_t contig_len,
                      int64_t start, int64_t end, int window_size, int window_offset, int32_t *out_gc,
                      cudaStream_t s, int64_t *launches) {
    int64_t npos = end - start;
    if (npos > 0) {
        k_gc<<<(unsigned)((npos + GC_BLK - 1) / GC_BLK), 256, 0, s>>>(ref, ref_start, ref_len, contig_len, start,
                                                                      end, window_size, window_offset, out_gc);
        LAUNCHED();
    }
    return cudaGetLastError();
}

