/*
 * TEST INFRASTRUCTURE ONLY -- CPU parity oracle for the pileup / binned statistics path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
 * build, link or call this file.  The product (pysamstats_b200) never does.
 *
 * What it restates (every function cites the reference file:line it follows; paths are
 * relative to /root/reference):
 *   - pysamstats/opt.pyx  : the 15 PileupStat and 5 BinnedStat accumulators, the drivers
 *     iter_pileup_default / iter_pileup_padded_chrom / iter_binned_chrom, rootmean / mean /
 *     gc_content.
 *   - pysam <0.16 / htslib 1.9-1.10 (install_requires at setup.py:56-57; NOT vendored under
 *     /root/reference): bam_plp_push / bam_plp_next / resolve_cigar2 / overlap_push /
 *     tweak_overlap_quality, restated from the published algorithm
 *     and anchored on the reference's call sites opt.pyx:1853-1854, :1945-1946, :2040.
 *
 * Parity pin: this oracle reproduces fixture/regression.npz (all 20 arrays) and the
 * known-answer tests of pysamstats/test/test_pileup.py:939-1102, test_binned.py:583-626
 * -- see tests/test_oracle_golden.py.
 *
 * Style: deliberately literal (a list of live reads walked per column, a name-keyed
 * table for the mate-overlap logic), single threaded, no cleverness.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define F_PAIRED 1
#define F_PROPER 2
#define F_UNMAP 4
#define F_MUNMAP 8
#define F_REVERSE 16
#define F_MREVERSE 32
#define F_SECONDARY 256
#define F_QCFAIL 512
#define F_DUP 1024

enum { C_M = 0, C_I, C_D, C_N, C_S, C_H, C_P, C_EQ, C_X };

typedef struct {
    int64_t n;
    const int32_t *pos;
    const int32_t *tlen;
    const int32_t *mtid;
    const int32_t *l_seq;
    const uint16_t *flag;
    const uint8_t *mapq;
    const uint32_t *cigar;
    const int64_t *cigar_off; /* n+1, in words */
    const uint8_t *seq4;
    const int64_t *seq_off; /* n+1, in bytes */
    const uint8_t *qual;
    const int64_t *qual_off; /* n+1, in bytes */
    const int64_t *name_id;  /* dense: equal QNAME <=> equal id */
    int64_t n_names;
    int32_t tid;
} orc_reads;

typedef struct {
    int32_t stat;     /* 0..14, order of pysamstats/pileup.py:150-166 */
    int32_t stepper;  /* 0 all, 1 nofilter, 2 samtools */
    int32_t max_depth;
    int32_t truncate;
    int32_t pad;
    int32_t min_mapq;
    int32_t min_baseq;
    int32_t no_del;
    int32_t no_dup;
    int32_t window_size;   /* coverage_gc */
    int32_t window_offset; /* coverage_gc, already defaulted */
    int32_t has_ref;
    int64_t start; /* normalised, zero based */
    int64_t end;
} orc_params;

/* ---------------------------------------------------------------- numerics */

/* opt.pyx:2159-2163 rootmean: int(np.around(sqrt(sqsum / count))); np.around is
 * round-half-to-even on float64 == rint() in the default rounding mode. */
static int32_t rootmean(uint64_t sqsum, int64_t count) {
    if (count > 0) return (int32_t)rint(sqrt((double)sqsum / (double)count));
    return 0;
}
/* opt.pyx:2166-2170 mean */
static int32_t mean_i(int64_t total, int64_t count) {
    if (count > 0) return (int32_t)rint((double)total / (double)count);
    return 0;
}
/* opt.pyx:1527-1539 gc_content: counts lower-cased g/c, divides by the fetched length */
static int32_t gc_percent(const uint8_t *ref, int64_t ref_len, int64_t ws, int64_t we) {
    if (ws < 0) ws = 0;
    if (we > ref_len) we = ref_len;
    if (we <= ws) return -1;
    int64_t gc = 0;
    for (int64_t i = ws; i < we; ++i) {
        uint8_t c = ref[i] | 0x20;
        if (c == 'g' || c == 'c') ++gc;
    }
    return (int32_t)rint((double)gc * 100. / (double)(we - ws));
}

/* ------------------------------------------------------- cigar / seq helpers */

static inline int cop(uint32_t w) { return (int)(w & 15u); }
static inline int64_t cln(uint32_t w) { return (int64_t)(w >> 4); }
static inline int is_match_op(int op) { return op == C_M || op == C_EQ || op == C_X; }

/* htslib bam_endpos */
static int64_t endpos_of(const orc_reads *R, int64_t i) {
    int64_t rl = 0;
    if (!(R->flag[i] & F_UNMAP)) {
        for (int64_t k = R->cigar_off[i]; k < R->cigar_off[i + 1]; ++k) {
            int op = cop(R->cigar[k]);
            if (op == C_M || op == C_D || op == C_N || op == C_EQ || op == C_X) rl += cln(R->cigar[k]);
        }
    }
    return (int64_t)R->pos[i] + (rl > 0 ? rl : 1);
}

/* opt.pyx:2133-2140 is_softclipped */
static int softclipped(const orc_reads *R, int64_t i) {
    for (int64_t k = R->cigar_off[i]; k < R->cigar_off[i + 1]; ++k)
        if (cop(R->cigar[k]) == C_S) return 1;
    return 0;
}

/* opt.pyx:2143-2156 get_seq_base: the 4-bit code at query offset k (0 if no sequence) */
static inline int seq_code(const orc_reads *R, int64_t i, int64_t k) {
    const uint8_t *p = R->seq4 + R->seq_off[i];
    return (p[k >> 1] >> (4 * (1 - (k & 1)))) & 0xf;
}
static const char NT16[] = "=ACMGRSVTWYHKDBN"; /* opt.pyx:72 */

/* bam_get_qual(b)[qpos] (opt.pyx:1902, :1166).  For a record without SEQ/QUAL (l_seq == 0) the
 * reference indexes past the empty quality array -- undefined behaviour; oracle and engine
 * both define that quality as 0. */
static inline int64_t qual_at(const orc_reads *R, const uint8_t *qual, int64_t i, int64_t qpos) {
    return qpos < R->l_seq[i] ? (int64_t)qual[R->qual_off[i] + qpos] : 0;
}

typedef struct {
    int is_del, is_refskip, indel;
    int64_t qpos;
} plp_event;

/* htslib resolve_cigar2, closed form per (read, column) -- SURVEY.md A.1 */
static void resolve(const orc_reads *R, int64_t i, int64_t pos, plp_event *e) {
    const uint32_t *cg = R->cigar + R->cigar_off[i];
    int64_t nc = R->cigar_off[i + 1] - R->cigar_off[i];
    int64_t x = R->pos[i], y = 0;
    e->is_del = e->is_refskip = e->indel = 0;
    e->qpos = 0;
    for (int64_t k = 0; k < nc; ++k) {
        int op = cop(cg[k]);
        int64_t l = cln(cg[k]);
        if (op == C_S || op == C_I) { y += l; continue; }
        if (op == C_H || op == C_P) continue;
        /* reference-consuming op */
        if (pos < x + l) {
            if (pos == x + l - 1 && k + 1 < nc) { /* peek the next operation */
                int op2 = cop(cg[k + 1]);
                int64_t l2 = cln(cg[k + 1]);
                if (op2 == C_D) e->indel = -(int)l2;
                else if (op2 == C_I) e->indel = (int)l2;
                else if (op2 == C_P && k + 2 < nc) {
                    int64_t l3 = 0;
                    for (int64_t j = k + 2; j < nc; ++j) {
                        op2 = cop(cg[j]);
                        l2 = cln(cg[j]);
                        if (op2 == C_I) l3 += l2;
                        else if (op2 == C_D || op2 == C_M || op2 == C_N || op2 == C_EQ || op2 == C_X) break;
                    }
                    if (l3 > 0) e->indel = (int)l3;
                }
            }
            if (is_match_op(op)) e->qpos = y + (pos - x);
            else { e->is_del = 1; e->qpos = y; e->is_refskip = (op == C_N); }
            return;
        }
        x += l;
        if (is_match_op(op)) y += l;
    }
    /* not reachable for pos < endpos */
    e->is_del = 1;
}

/* --------------------------------------- mate-overlap tweak */

typedef struct { int64_t ci, icig, iseq, iref; } cur_t;

static int cur_set(const uint32_t *cg, int64_t nc, int64_t want, cur_t *c) {
    c->ci = c->icig = c->iseq = c->iref = 0;
    if (want < 0) return 0;
    while (c->ci < nc) {
        int op = cop(cg[c->ci]);
        int64_t n = cln(cg[c->ci]);
        if (op == C_S) { c->ci++; c->iseq += n; c->icig = 0; }
        else if (op == C_H || op == C_P) { c->ci++; c->icig = 0; }
        else if (is_match_op(op)) {
            want -= n;
            if (want < 0) { c->icig = n + want; c->iseq += c->icig; c->iref += c->icig; return 1; }
            c->ci++; c->iseq += n; c->iref += n; c->icig = 0;
        } else if (op == C_I) { c->ci++; c->iseq += n; c->icig = 0; }
        else { /* D, N */
            want -= n;
            if (want < 0) want = 0;
            c->ci++; c->iref += n; c->icig = 0;
        }
    }
    return 0;
}

static int cur_next(const uint32_t *cg, int64_t nc, cur_t *c) {
    while (c->ci < nc) {
        int op = cop(cg[c->ci]);
        int64_t n = cln(cg[c->ci]);
        if (is_match_op(op)) {
            if (c->icig >= n - 1) { c->icig = 0; c->ci++; continue; } /* old-htslib reset to 0 */
            c->iseq++; c->icig++; c->iref++;
            return 1;
        } else if (op == C_D || op == C_N) { c->ci++; c->iref += n; c->icig = 0; }
        else if (op == C_I || op == C_S) { c->ci++; c->iseq += n; c->icig = 0; }
        else { c->ci++; c->icig = 0; }
    }
    c->iseq = c->iref = -1;
    return 0;
}

static void tweak(const orc_reads *R, uint8_t *qual, int64_t a, int64_t b) {
    const uint32_t *ca = R->cigar + R->cigar_off[a], *cb = R->cigar + R->cigar_off[b];
    int64_t na = R->cigar_off[a + 1] - R->cigar_off[a], nb = R->cigar_off[b + 1] - R->cigar_off[b];
    int64_t apos = R->pos[a], bpos = R->pos[b];
    uint8_t *qa = qual + R->qual_off[a], *qb = qual + R->qual_off[b];
    int64_t iref = bpos;
    cur_t A, B;
    if (!cur_set(ca, na, iref - apos, &A)) return;
    if (!cur_set(cb, nb, iref - bpos, &B)) return;
    int ra = 1, rb = 1;
    for (;;) {
        while (A.iref >= 0 && A.iref < iref - apos) ra = cur_next(ca, na, &A);
        if (!ra) return;
        if (iref < A.iref + apos) iref = A.iref + apos;
        while (B.iref >= 0 && B.iref < iref - bpos) rb = cur_next(cb, nb, &B);
        if (!rb) return;
        if (iref < B.iref + bpos) iref = B.iref + bpos;
        iref++;
        if (A.iref + apos != B.iref + bpos) continue;
        int va = qa[A.iseq], vb = qb[B.iseq];
        if (seq_code(R, a, A.iseq) == seq_code(R, b, B.iseq)) {
            int s = va + vb;
            qa[A.iseq] = (uint8_t)(s > 200 ? 200 : s);
            qb[B.iseq] = 0;
        } else if (va >= vb) {
            qa[A.iseq] = (uint8_t)(va * 4 / 5); /* floor(0.8*q) */
            qb[B.iseq] = 0;
        } else {
            qb[B.iseq] = (uint8_t)(vb * 4 / 5);
            qa[A.iseq] = 0;
        }
    }
}

/* ------------------------------------------------------------ accumulators */

typedef struct { int32_t all, fwd, rev, pp, pp_fwd, pp_rev; } cps_t; /* opt.pyx:147-168 */
static inline void cps_incr(cps_t *c, int rev, int pp) {
    c->all++;
    if (rev) { c->rev++; if (pp) { c->pp++; c->pp_rev++; } }
    else { c->fwd++; if (pp) { c->pp++; c->pp_fwd++; } }
}

typedef struct { /* opt.pyx:679-727 TlenHelper */
    int64_t n; double m, m2; int64_t s, sq;
} tl_t;
static inline void tl_update(tl_t *t, int64_t x) {
    t->n++; t->s += x; t->sq += x * x;
    double d = (double)x - t->m;
    t->m += d / (double)t->n;
    double d2 = (double)x - t->m;
    t->m2 += d * d2;
}
static inline int32_t tl_std(const tl_t *t) {
    if (t->n < 2) return 0;
    return (int32_t)rint(sqrt(t->m2 / (double)(t->n - 1)));
}

typedef struct { int32_t n, nz; uint64_t mx, sq; } mq_t; /* opt.pyx:912-937 */
static inline void mq_update(mq_t *m, uint64_t q) {
    m->n++; if (q == 0) m->nz++; if (q > m->mx) m->mx = q; m->sq += q * q;
}

typedef struct { int32_t n, n_nodel; uint64_t sq; } bq_t; /* opt.pyx:1093-1115 */
static inline void bq_update(bq_t *b, int64_t sq) {
    b->n++; if (sq >= 0) { b->n_nodel++; b->sq += (uint64_t)sq; }
}
/* opt.pyx:1188-1226: index 0 all,1 fwd,2 rev,3 pp,4 pp_fwd,5 pp_rev */
static inline void bq6_update(bq_t *h, int64_t sq, int pp, int rev) {
    bq_update(&h[0], sq);
    bq_update(rev ? &h[2] : &h[1], sq);
    if (pp) { bq_update(&h[3], sq); bq_update(rev ? &h[5] : &h[4], sq); }
}

typedef struct {
    cps_t c[10];   /* coverage*, variation*: reads, matches, mismatches, deletions, insertions, A, C, T, G, N */
    int32_t ext[8][3]; /* coverage_ext(_strand): class x {all,fwd,rev} */
    tl_t tl[6];    /* all,fwd,rev,pp,pp_fwd,pp_rev */
    int32_t n_all, n_fwd, n_rev;
    mq_t mq[6];
    bq_t bq[3][6]; /* all / matches / mismatches */
} acc_t;

static const int NFIELDS[15] = {2, 6, 8, 24, 20, 60, 9, 27, 8, 24, 4, 12, 12, 36, 3};

int32_t orc_nfields(int32_t stat) { return (stat >= 0 && stat < 15) ? NFIELDS[stat] : -1; }

/* the per-event dispatch: one recv() per stat class of opt.pyx */
static void recv(int stat, acc_t *A, const orc_reads *R, const uint8_t *qual, int64_t i,
                 const plp_event *e, int refcode_char /* upper-cased ref char or -1 */) {
    uint32_t flag = R->flag[i];
    int rev = (flag & F_REVERSE) != 0, pp = (flag & F_PROPER) != 0;
    switch (stat) {
    case 0: case 1: case 14: /* Coverage opt.pyx:119-127, CoverageStrand :184-196, CoverageGC :1488-1496 */
        cps_incr(&A->c[0], rev, pp);
        break;
    case 2: case 3: { /* CoverageExt opt.pyx:245-279, CoverageExtStrand :357-393 */
        int mun = (flag & F_MUNMAP) != 0, mrev = (flag & F_MREVERSE) != 0;
        int32_t tlen = R->tlen[i];
        int cls[8] = {1, pp, 0, 0, 0, 0, softclipped(R, i), (flag & F_DUP) != 0};
        if (mun) cls[2] = 1;
        else if (R->tid != R->mtid[i]) cls[3] = 1;
        else if ((rev && mrev) || (!rev && !mrev)) cls[4] = 1;
        else if ((rev && tlen > 0) || (!rev && tlen < 0)) cls[5] = 1;
        for (int k = 0; k < 8; ++k)
            if (cls[k]) { A->ext[k][0]++; A->ext[k][rev ? 2 : 1]++; }
        break;
    }
    case 4: case 5: { /* Variation opt.pyx:474-506, VariationStrand :583-616 */
        cps_incr(&A->c[0], rev, pp);
        if (e->is_del) {
            if (!e->is_refskip) cps_incr(&A->c[3], rev, pp);
        } else {
            int base = R->l_seq[i] ? NT16[seq_code(R, i, e->qpos)] : -2;
            if (base == 'A') cps_incr(&A->c[5], rev, pp);
            else if (base == 'T') cps_incr(&A->c[7], rev, pp);
            else if (base == 'C') cps_incr(&A->c[6], rev, pp);
            else if (base == 'G') cps_incr(&A->c[8], rev, pp);
            else if (base == 'N') cps_incr(&A->c[9], rev, pp);
            if (e->indel > 0) cps_incr(&A->c[4], rev, pp);
            if (base == refcode_char) cps_incr(&A->c[1], rev, pp);
            else cps_incr(&A->c[2], rev, pp);
        }
        break;
    }
    case 6: case 7: { /* Tlen opt.pyx:748-769, TlenStrand :829-865 */
        int mun = (flag & F_MUNMAP) != 0;
        int other = R->tid != R->mtid[i];
        int64_t tlen = R->tlen[i];
        A->n_all++;
        if (rev) A->n_rev++; else A->n_fwd++;
        if (!mun && !other) { tl_update(&A->tl[0], tlen); tl_update(&A->tl[rev ? 2 : 1], tlen); }
        if (pp) { tl_update(&A->tl[3], tlen); tl_update(&A->tl[rev ? 5 : 4], tlen); }
        break;
    }
    case 8: case 9: { /* Mapq opt.pyx:956-971, MapqStrand :1024-1049 */
        uint64_t q = R->mapq[i];
        mq_update(&A->mq[0], q);
        mq_update(&A->mq[rev ? 2 : 1], q);
        if (pp) { mq_update(&A->mq[3], q); mq_update(&A->mq[rev ? 5 : 4], q); }
        break;
    }
    case 10: case 11: { /* Baseq opt.pyx:1153-1169, BaseqStrand :1242-1260 */
        int64_t sq = -1;
        if (!e->is_del) { int64_t q = qual_at(R, qual, i, e->qpos); sq = q * q; }
        bq6_update(A->bq[0], sq, pp, rev);
        break;
    }
    case 12: case 13: { /* BaseqExt opt.pyx:1310-1335, BaseqExtStrand :1385-1410 */
        int64_t sq = -1;
        if (!e->is_del) { int64_t q = qual_at(R, qual, i, e->qpos); sq = q * q; }
        bq6_update(A->bq[0], sq, pp, rev);
        if (!e->is_del) {
            int base = R->l_seq[i] ? NT16[seq_code(R, i, e->qpos)] : -2;
            bq6_update(A->bq[base == refcode_char ? 1 : 2], sq, pp, rev);
        }
        break;
    }
    default: break;
    }
}

#define PUT6(c) do { *o++ = (c).all; *o++ = (c).fwd; *o++ = (c).rev; *o++ = (c).pp; *o++ = (c).pp_fwd; *o++ = (c).pp_rev; } while (0)
#define PUT2(c) do { *o++ = (c).all; *o++ = (c).pp; } while (0)

/* the rec() of each stat class: field order = pysamstats/config.py dtype lists (minus
 * chrom, pos, ref); resets the accumulator like every rec() does. */
static void rec(int stat, acc_t *A, int32_t *o, int32_t gc) {
    static const int six[6] = {0, 1, 2, 3, 4, 5};
    switch (stat) {
    case 0: PUT2(A->c[0]); break;                              /* opt.pyx:129-138 */
    case 1: PUT6(A->c[0]); break;                              /* opt.pyx:198-211 */
    case 2: for (int k = 0; k < 8; ++k) *o++ = A->ext[k][0]; break; /* opt.pyx:281-296 */
    case 3: for (int k = 0; k < 8; ++k) { *o++ = A->ext[k][0]; *o++ = A->ext[k][1]; *o++ = A->ext[k][2]; } break; /* :395-426 */
    case 4: for (int k = 0; k < 10; ++k) PUT2(A->c[k]); break; /* opt.pyx:508-535 */
    case 5: for (int k = 0; k < 10; ++k) PUT6(A->c[k]); break; /* opt.pyx:618-670 */
    case 6: /* opt.pyx:771-787 */
        *o++ = A->n_all; *o++ = (int32_t)A->tl[0].n; *o++ = (int32_t)A->tl[3].n;
        *o++ = mean_i(A->tl[0].s, A->tl[0].n); *o++ = mean_i(A->tl[3].s, A->tl[3].n);
        *o++ = rootmean((uint64_t)A->tl[0].sq, A->tl[0].n); *o++ = rootmean((uint64_t)A->tl[3].sq, A->tl[3].n);
        *o++ = tl_std(&A->tl[0]); *o++ = tl_std(&A->tl[3]);
        break;
    case 7: /* opt.pyx:867-903 */
        *o++ = A->n_all; *o++ = A->n_fwd; *o++ = A->n_rev;
        for (int k = 0; k < 6; ++k) *o++ = (int32_t)A->tl[six[k]].n;
        for (int k = 0; k < 6; ++k) *o++ = mean_i(A->tl[k].s, A->tl[k].n);
        for (int k = 0; k < 6; ++k) *o++ = rootmean((uint64_t)A->tl[k].sq, A->tl[k].n);
        for (int k = 0; k < 6; ++k) *o++ = tl_std(&A->tl[k]);
        break;
    case 8: /* opt.pyx:973-988 */
        *o++ = A->mq[0].n; *o++ = A->mq[3].n; *o++ = A->mq[0].nz; *o++ = A->mq[3].nz;
        *o++ = rootmean(A->mq[0].sq, A->mq[0].n); *o++ = rootmean(A->mq[3].sq, A->mq[3].n);
        *o++ = (int32_t)A->mq[0].mx; *o++ = (int32_t)A->mq[3].mx;
        break;
    case 9: /* opt.pyx:1051-1084 */
        for (int k = 0; k < 6; ++k) *o++ = A->mq[k].n;
        for (int k = 0; k < 6; ++k) *o++ = A->mq[k].nz;
        for (int k = 0; k < 6; ++k) *o++ = rootmean(A->mq[k].sq, A->mq[k].n);
        for (int k = 0; k < 6; ++k) *o++ = (int32_t)A->mq[k].mx;
        break;
    case 10: /* opt.pyx:1171-1184 */
        *o++ = A->bq[0][0].n; *o++ = A->bq[0][3].n;
        *o++ = rootmean(A->bq[0][0].sq, A->bq[0][0].n_nodel); *o++ = rootmean(A->bq[0][3].sq, A->bq[0][3].n_nodel);
        break;
    case 11: /* opt.pyx:1262-1283 */
        for (int k = 0; k < 6; ++k) *o++ = A->bq[0][k].n;
        for (int k = 0; k < 6; ++k) *o++ = rootmean(A->bq[0][k].sq, A->bq[0][k].n_nodel);
        break;
    case 12: /* opt.pyx:1337-1363 */
        for (int g = 0; g < 3; ++g) { *o++ = A->bq[g][0].n; *o++ = A->bq[g][3].n; }
        for (int g = 0; g < 3; ++g) { *o++ = rootmean(A->bq[g][0].sq, A->bq[g][0].n_nodel); *o++ = rootmean(A->bq[g][3].sq, A->bq[g][3].n_nodel); }
        break;
    case 13: /* opt.pyx:1412-1462 */
        for (int g = 0; g < 3; ++g) for (int k = 0; k < 6; ++k) *o++ = A->bq[g][k].n;
        for (int g = 0; g < 3; ++g) for (int k = 0; k < 6; ++k) *o++ = rootmean(A->bq[g][k].sq, A->bq[g][k].n_nodel);
        break;
    case 14: *o++ = gc; PUT2(A->c[0]); break; /* opt.pyx:1498-1519; config.py dtype_coverage_gc order */
    default: break;
    }
    memset(A, 0, sizeof(*A));
}

/* ------------------------------------------------------------ pileup driver */

typedef struct {
    const orc_reads *R;
    const orc_params *P;
    const uint8_t *ref; /* contig sequence as in the FASTA (any case), may be NULL */
    int64_t ref_len;
    uint8_t *qual;      /* private, tweaked copy */
    int64_t *out_pos;
    int32_t *out_fields;
    uint8_t *out_ref;
    int64_t cap, n_rows;
    int nf;
    acc_t acc;
} drv_t;

/* opt.pyx:1987-1994 get_refbase -> upper-cased char, or -1 when fetch() is empty / no fasta */
static int refchar(const drv_t *D, int64_t pos) {
    if (!D->ref || pos < 0 || pos >= D->ref_len) return -1;
    int c = D->ref[pos];
    if (c >= 'a' && c <= 'z') c -= 32;
    return c;
}

static int32_t gc_at(const drv_t *D, int64_t pos) {
    if (D->P->stat != 14) return 0;
    int64_t ws = pos - D->P->window_offset; /* opt.pyx:1501-1509 */
    int64_t we = ws + D->P->window_size;
    return gc_percent(D->ref, D->ref_len, ws, we);
}

static void emit_row(drv_t *D, int64_t pos) {
    if (D->n_rows < D->cap) {
        D->out_pos[D->n_rows] = pos;
        rec(D->P->stat, &D->acc, D->out_fields + D->n_rows * D->nf, gc_at(D, pos));
        int c = refchar(D, pos);
        if (D->out_ref) D->out_ref[D->n_rows] = (uint8_t)(c < 0 ? 0 : c);
    } else {
        memset(&D->acc, 0, sizeof(D->acc));
    }
    D->n_rows++;
}

/* opt.pyx:1866-1919 stat_pileup: filters, recv per read in the column, then rec */
static void do_column(drv_t *D, int64_t pos, const int64_t *col, int64_t ncol) {
    const orc_reads *R = D->R;
    const orc_params *P = D->P;
    int rc = refchar(D, pos);
    for (int64_t j = 0; j < ncol; ++j) {
        int64_t i = col[j];
        plp_event e;
        resolve(R, i, pos, &e);
        if (P->min_mapq > 0 && R->mapq[i] < P->min_mapq) continue;
        if (P->min_baseq > 0 && !e.is_del && qual_at(R, D->qual, i, e.qpos) < P->min_baseq) continue;
        if (P->no_del && e.is_del) continue;
        if (P->no_dup && (R->flag[i] & F_DUP)) continue;
        recv(P->stat, &D->acc, R, D->qual, i, &e, rc);
    }
    emit_row(D, pos);
}

/*
 * Pileup statistics for the reads of one contig region (already selected by the region
 * rule of bamread.region_reads, file order).  Returns the number of rows; if it exceeds
 * `cap` the caller retries with a larger buffer.
 *
 * Follows opt.pyx:1846-1863 (iter_pileup_default) / :1936-1984 (iter_pileup_padded_chrom)
 * on top of a literal bam_plp_push / bam_plp_next loop.
 */
int64_t orc_pileup(const orc_reads *R, const orc_params *P, const uint8_t *ref, int64_t ref_len,
                   int64_t *out_pos, int32_t *out_fields, uint8_t *out_ref, int64_t cap) {
    drv_t D;
    memset(&D, 0, sizeof(D));
    D.R = R; D.P = P; D.ref = P->has_ref ? ref : NULL; D.ref_len = ref_len;
    D.out_pos = out_pos; D.out_fields = out_fields; D.out_ref = out_ref; D.cap = cap;
    D.nf = NFIELDS[P->stat];
    int64_t n = R->n;
    int64_t qbytes = n ? R->qual_off[n] : 0;
    D.qual = (uint8_t *)malloc((size_t)(qbytes > 0 ? qbytes : 1));
    if (qbytes > 0) memcpy(D.qual, R->qual, (size_t)qbytes);
    int64_t *endp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t *live = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1)); /* list of nodes, list order */
    int64_t *col = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t *stored = (int64_t *)malloc(sizeof(int64_t) * (size_t)(R->n_names + 1)); /* overlap hash */
    for (int64_t k = 0; k < R->n_names; ++k) stored[k] = -1;
    int64_t nlive = 0;
    int64_t maxcnt = P->max_depth ? P->max_depth : 8000; /* A.0: falsy -> htslib default */
    int64_t it_pos = 0, max_pos = -1;
    int it_on_tid = (R->tid == 0); /* iterator starts at (tid 0, pos 0) */
    int64_t curpos = P->start;     /* pad cursor, opt.pyx:1949 */
    int stop = 0;

    for (int64_t i = 0; i <= n && !stop; ++i) {
        int eof = (i == n);
        if (!eof) {
            uint32_t flag = R->flag[i];
            /* stepper, applied before the read reaches the pileup engine (A.0) */
            if (P->stepper == 0 || P->stepper == 2) {
                if (flag & (F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP)) continue;
                if (P->stepper == 2 && (flag & F_PAIRED) && !(flag & F_PROPER)) continue;
            }
            /* bam_plp_push */
            if (R->tid < 0 || (flag & F_UNMAP)) { stored[R->name_id[i]] = -1; continue; }
            if (it_on_tid && it_pos == R->pos[i] && nlive >= maxcnt) { stored[R->name_id[i]] = -1; continue; }
            endp[i] = endpos_of(R, i);
            /* overlap_push */
            if (!(flag & F_MUNMAP) && (flag & F_PROPER) &&
                llabs((long long)R->tlen[i]) < 2LL * R->l_seq[i]) {
                int64_t nm = R->name_id[i];
                if (stored[nm] < 0) stored[nm] = i;
                else { tweak(R, D.qual, stored[nm], i); stored[nm] = -1; }
            }
            max_pos = R->pos[i];
            if (!it_on_tid || endp[i] > it_pos) live[nlive++] = i;
        }
        /* bam_plp_next until it has no column to give */
        while (eof ? nlive > 0 : (!it_on_tid || max_pos > it_pos)) {
            int64_t ncol = 0, w = 0;
            if (it_on_tid) {
                for (int64_t j = 0; j < nlive; ++j) {
                    int64_t r = live[j];
                    if (endp[r] <= it_pos) { stored[R->name_id[r]] = -1; continue; } /* expire */
                    live[w++] = r;
                    if (R->pos[r] <= it_pos) col[ncol++] = r;
                }
                nlive = w;
            }
            int64_t this_pos = it_pos;
            /* advance the iterator */
            if (!it_on_tid) { it_on_tid = 1; it_pos = R->pos[live[0]]; }
            else if (nlive > 0 && it_pos < R->pos[live[0]]) it_pos = R->pos[live[0]];
            else ++it_pos;
            if (ncol == 0) continue;
            /* pysam IteratorColumnRegion truncate (A.0) */
            if (P->truncate) {
                if (this_pos < P->start) continue;
                if (this_pos >= P->end) { stop = 1; break; }
            }
            if (P->pad) { /* opt.pyx:1955-1964 */
                while (curpos < this_pos) { emit_row(&D, curpos); curpos++; }
            }
            do_column(&D, this_pos, col, ncol);
            if (P->pad) curpos = this_pos + 1; /* opt.pyx:1973 */
        }
    }
    if (P->pad) { /* opt.pyx:1975-1984 */
        while (curpos < P->end) { emit_row(&D, curpos); curpos++; }
    }
    free(D.qual); free(endp); free(live); free(col); free(stored);
    return D.n_rows;
}

/* ------------------------------------------------------------ binned driver */

static const int NFIELDS_BINNED[5] = {3, 9, 3, 11, 6};
int32_t orc_nfields_binned(int32_t stat) { return (stat >= 0 && stat < 5) ? NFIELDS_BINNED[stat] : -1; }

typedef struct {
    int32_t cov[8];
    int32_t n_all, n_mapq0; uint64_t mapq_sq;
    int32_t ops[9];
    int32_t n_pp; int64_t tl_s, tl_pp_s, tl_sq, tl_pp_sq;
} bacc_t;

static void brecv(int stat, bacc_t *A, const orc_reads *R, int64_t i) {
    uint32_t flag = R->flag[i];
    if (flag & F_UNMAP) return;
    int rev = (flag & F_REVERSE) != 0, pp = (flag & F_PROPER) != 0;
    switch (stat) {
    case 0: A->cov[0]++; if (pp) A->cov[1]++; break; /* opt.pyx:1574-1584 */
    case 1: { /* opt.pyx:1629-1662 */
        int mun = (flag & F_MUNMAP) != 0, mrev = (flag & F_MREVERSE) != 0;
        int32_t tlen = R->tlen[i];
        A->cov[0]++;
        if (flag & F_DUP) A->cov[7]++;
        if (pp) A->cov[1]++;
        if (mun) A->cov[2]++;
        else if (R->tid != R->mtid[i]) A->cov[3]++;
        else if ((rev && mrev) || (!rev && !mrev)) A->cov[4]++;
        else if ((rev && tlen > 0) || (!rev && tlen < 0)) A->cov[5]++;
        if (softclipped(R, i)) A->cov[6]++;
        break;
    }
    case 2: { /* opt.pyx:1690-1703 */
        uint64_t q = R->mapq[i];
        A->n_all++; A->mapq_sq += q * q; if (q == 0) A->n_mapq0++;
        break;
    }
    case 3: /* opt.pyx:1733-1764 */
        for (int64_t k = R->cigar_off[i]; k < R->cigar_off[i + 1]; ++k) {
            int op = cop(R->cigar[k]);
            if (op <= C_X) A->ops[op] += (int32_t)cln(R->cigar[k]);
        }
        A->n_all++;
        break;
    case 4: { /* opt.pyx:1801-1819 */
        int64_t t = R->tlen[i];
        A->n_all++; A->tl_s += t; A->tl_sq += t * t;
        if (pp) { A->n_pp++; A->tl_pp_s += t; A->tl_pp_sq += t * t; }
        break;
    }
    default: break;
    }
}

static void brec(int stat, bacc_t *A, int32_t *o, int32_t gc) {
    switch (stat) {
    case 0: *o++ = gc; *o++ = A->cov[0]; *o++ = A->cov[1]; break;          /* opt.pyx:1558-1572 */
    case 1: *o++ = gc; for (int k = 0; k < 8; ++k) *o++ = A->cov[k]; break; /* opt.pyx:1604-1627 */
    case 2: *o++ = A->n_all; *o++ = A->n_mapq0; *o++ = rootmean(A->mapq_sq, A->n_all); break; /* :1678-1688 */
    case 3: /* opt.pyx:1719-1731; config.py dtype_alignment_binned order */
        *o++ = A->n_all;
        *o++ = A->ops[C_M] + A->ops[C_I] + A->ops[C_S] + A->ops[C_EQ] + A->ops[C_X];
        for (int k = 0; k < 9; ++k) *o++ = A->ops[k];
        break;
    case 4: /* opt.pyx:1785-1799 */
        *o++ = A->n_all; *o++ = A->n_pp;
        *o++ = mean_i(A->tl_s, A->n_all); *o++ = mean_i(A->tl_pp_s, A->n_pp);
        *o++ = rootmean((uint64_t)A->tl_sq, A->n_all); *o++ = rootmean((uint64_t)A->tl_pp_sq, A->n_pp);
        break;
    default: break;
    }
    memset(A, 0, sizeof(*A));
}

/*
 * opt.pyx:2026-2102 iter_binned_chrom, literally: roll the window while pos > bin_end,
 * filters after the roll-over, last bin, then empty bins up to `end`.
 * out_pos receives bin_start + window_offset (zero based).
 */
int64_t orc_binned(const orc_reads *R, int32_t stat, int64_t start, int64_t end, int32_t window_size,
                   int32_t window_offset, int32_t min_mapq, int32_t no_dup, const uint8_t *ref,
                   int64_t ref_len, int64_t *out_pos, int32_t *out_fields, int64_t cap) {
    int nf = NFIELDS_BINNED[stat];
    bacc_t A;
    memset(&A, 0, sizeof(A));
    int64_t rows = 0;
    int64_t bin_start = start, bin_end = start + window_size;
    int needs_gc = (stat == 0 || stat == 1);
#define EMIT_BIN() do { \
        if (rows < cap) { \
            int32_t gc = needs_gc ? gc_percent(ref, ref_len, bin_start, bin_end) : 0; \
            out_pos[rows] = bin_start + window_offset; \
            brec(stat, &A, out_fields + rows * nf, gc); \
        } else memset(&A, 0, sizeof(A)); \
        rows++; } while (0)
    for (int64_t i = 0; i < R->n; ++i) {
        while (R->pos[i] > bin_end) {
            EMIT_BIN();
            bin_start = bin_end;
            bin_end = bin_start + window_size;
        }
        if (min_mapq > 0 && R->mapq[i] < min_mapq) continue;
        if (no_dup && (R->flag[i] & F_DUP)) continue;
        brecv(stat, &A, R, i);
    }
    EMIT_BIN();
    while (bin_end < end) {
        bin_start = bin_end;
        bin_end = bin_start + window_size;
        EMIT_BIN();
    }
#undef EMIT_BIN
    return rows;
}
