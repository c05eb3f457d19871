// C ABI of libpss_b200 (include/pss_b200.h): argument checking, workspace management,
// kernel sequencing and the host-buffer (streaming) entry points.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "pss_expand.h"
#include "pss_internal.cuh"

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct pss_ctx {
    int device = 0;
    char err[512] = {0};
    int64_t launches = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;   // moved by the *_host entry points
    DevBuf ri, rng, qual, pair, next, small, tdesc, refpack, bwork, redo, lowq;
    int lowq_minq = -1;     // min_baseq the marks in `lowq` were made for (-1: none; reset with every prepared batch)
    int64_t qual_bias = 0;  // ctx->qual holds bytes [qual_bias, qual_bias + n) of the batch's qual array
    // batch prepared by pss_prepare()
    bool prepared = false;
    pss_reads prep_reads;
    int prep_stepper = 0, prep_max_depth = 0, prep_min_mapq = 0, prep_no_dup = 0, prep_qual = 0;
    int64_t desc_start = -1, desc_end = -1;   // position range the tile descriptors in tdesc describe
    DevBuf hb[20];          // device copies of host batches (*_host entry points)
    DevBuf hb2[20];         // second set: pss_pileup_host_stream uploads chunk k + 1 while chunk k is computed
    cudaStream_t up_stream = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    DevBuf outbuf[2], maskbuf[2];
    // narrow (8- / 16-bit) result transport of the host entry points: 3 device buffers, 3 pinned staging buffers
    DevBuf nbuf[3], nmask[3];
    void *hstage[3] = {nullptr, nullptr, nullptr};
    DevBuf nflag;                       // per-buffer "a count needs more than 8 bits" flags
    int32_t *hflag = nullptr;           // pinned mirror
    size_t hstage_bytes[3] = {0, 0, 0};
    cudaEvent_t nev_done[3] = {nullptr, nullptr, nullptr}, nev_copied[3] = {nullptr, nullptr, nullptr};
    pss_pool *pool = nullptr;
    SmallState *small_host = nullptr;   // pinned mirror of the device scalars
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    // dominant-kernel timing
    int timing = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_stat;           // statistic id of event pair i / 2
    size_t ev_used = 0;
};

static int fail(pss_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, PSS_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

static int ensure(pss_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.bytes >= bytes) return PSS_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.bytes = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, PSS_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    b.bytes = want;
    return PSS_OK;
}

// ---------------------------------------------------------------------------- shapes

static const int32_t kPileupFields[PSS_N_PILEUP_STATS] = {2, 6, 8, 24, 20, 60, 9, 27, 8, 24, 4, 12, 12, 36, 3};
static const int32_t kBinnedFields[PSS_N_BINNED_STATS] = {3, 9, 3, 11, 6};

extern "C" int32_t pss_pileup_nfields(int32_t stat) {
    return (stat >= 0 && stat < PSS_N_PILEUP_STATS) ? kPileupFields[stat] : -1;
}
extern "C" int32_t pss_binned_nfields(int32_t stat) {
    return (stat >= 0 && stat < PSS_N_BINNED_STATS) ? kBinnedFields[stat] : -1;
}
extern "C" int32_t pss_pileup_needs_ref(int32_t stat) {
    return stat == PSS_VARIATION || stat == PSS_VARIATION_STRAND || stat == PSS_BASEQ_EXT ||
           stat == PSS_BASEQ_EXT_STRAND || stat == PSS_COVERAGE_GC;
}
extern "C" int32_t pss_pileup_needs_qual(int32_t stat, int32_t min_baseq) {
    return (stat >= PSS_BASEQ && stat <= PSS_BASEQ_EXT_STRAND) || min_baseq > 0;
}
extern "C" int32_t pss_tile_positions(void) { return PSS_TTILE; }
extern "C" int64_t pss_binned_nbins(int64_t start, int64_t end, int32_t window_size) {
    if (window_size <= 0) return -1;
    int64_t span = end - start;
    int64_t n = span > 0 ? (span + window_size - 1) / window_size : 0;
    return n < 1 ? 1 : n;
}

// ------------------------------------------------------------------------ life cycle

extern "C" const char *pss_version(void) { return "pss_b200 0.1.0 (sm_100a)"; }

extern "C" int pss_create(int device, pss_ctx **out) {
    if (!out) return PSS_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return PSS_ECUDA;
    }
    pss_ctx *ctx = new pss_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || ensure(ctx, ctx->small, sizeof(SmallState)) != PSS_OK ||
        cudaMallocHost((void **)&ctx->small_host, sizeof(SmallState)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return PSS_ECUDA;
    }
    memset(ctx->small_host, 0, sizeof(SmallState));
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming);
    }
    *out = ctx;
    return PSS_OK;
}

extern "C" void pss_destroy(pss_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf *all[] = {&ctx->ri, &ctx->rng, &ctx->qual, &ctx->pair, &ctx->next, &ctx->small, &ctx->tdesc, &ctx->refpack, &ctx->bwork, &ctx->redo, &ctx->lowq,
                     &ctx->outbuf[0], &ctx->outbuf[1], &ctx->maskbuf[0], &ctx->maskbuf[1]};
    for (DevBuf *b : all) if (b->p) cudaFree(b->p);
    for (DevBuf &b : ctx->hb) if (b.p) cudaFree(b.p);
    for (DevBuf &b : ctx->hb2) if (b.p) cudaFree(b.p);
    if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
        if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
    }
    if (ctx->small_host) cudaFreeHost(ctx->small_host);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 3; ++i) {
        if (ctx->nbuf[i].p) cudaFree(ctx->nbuf[i].p);
        if (ctx->nmask[i].p) cudaFree(ctx->nmask[i].p);
        if (ctx->hstage[i]) cudaFreeHost(ctx->hstage[i]);
        if (ctx->nev_done[i]) cudaEventDestroy(ctx->nev_done[i]);
        if (ctx->nev_copied[i]) cudaEventDestroy(ctx->nev_copied[i]);
    }
    if (ctx->nflag.p) cudaFree(ctx->nflag.p);
    if (ctx->hflag) cudaFreeHost(ctx->hflag);
    pss_pool_destroy(ctx->pool);
    delete ctx;
}

extern "C" const char *pss_last_error(const pss_ctx *ctx) { return ctx ? ctx->err : "no context"; }
extern "C" int64_t pss_launch_count(const pss_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void pss_transfer_bytes(const pss_ctx *ctx, int64_t *h2d, int64_t *d2h) {
    if (h2d) *h2d = ctx ? ctx->h2d_bytes : 0;
    if (d2h) *d2h = ctx ? ctx->d2h_bytes : 0;
}
extern "C" void pss_set_kernel_timing(pss_ctx *ctx, int enabled) { if (ctx) ctx->timing = enabled; }

extern "C" double pss_pileup_kernel_ms(pss_ctx *ctx, int32_t stat, int64_t *n_launches) {
    if (n_launches) *n_launches = 0;
    if (!ctx || ctx->ev_used == 0) return -1.0;
    cudaSetDevice(ctx->device);
    double sum = 0;
    int64_t n = 0;
    size_t kept = 0;
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        if (stat >= 0 && ctx->ev_stat[i / 2] != stat) {   // keep pairs of other statistics
            std::swap(ctx->ev_pool[kept], ctx->ev_pool[i]);
            std::swap(ctx->ev_pool[kept + 1], ctx->ev_pool[i + 1]);
            ctx->ev_stat[kept / 2] = ctx->ev_stat[i / 2];
            kept += 2;
            continue;
        }
        float ms = 0;
        if (cudaEventSynchronize(ctx->ev_pool[i + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]) == cudaSuccess) {
            sum += ms; ++n;
        } else cudaGetLastError();
    }
    ctx->ev_used = kept;
    if (n_launches) *n_launches = n;
    return n ? sum / (double)n : -1.0;
}

extern "C" int pss_synchronize(pss_ctx *ctx, void *stream) {
    if (!ctx) return PSS_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->small_host, ctx->small.p, sizeof(SmallState), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (ctx->small_host->unsorted) return fail(ctx, PSS_EUNSORTED, "read batch is not coordinate sorted");
    if (ctx->small_host->toolong)
        return fail(ctx, PSS_ETOOLONG, "a read is too long for the kernels' staging buffers; set "
                                       "PSS_FLAG_LONG_READS in pss_pileup_params.flags");
    return PSS_OK;
}

// -------------------------------------------------------------------- pileup (device)

static int check_reads(pss_ctx *ctx, const pss_reads *R, bool need_qual) {
    if (!R) return fail(ctx, PSS_EINVAL, "reads is NULL");
    if (R->n_reads < 0 || R->n_reads > 0x7fffffff) return fail(ctx, PSS_EINVAL, "n_reads out of range");
    if (R->n_reads > 0) {
        if (!R->pos || !R->flag || !R->mapq || !R->tlen || !R->mtid || !R->l_seq || !R->cigar_off ||
            !R->cigar || !R->seq_off || !R->seq4)
            return fail(ctx, PSS_EINVAL, "a required read array is NULL");
        if (need_qual && (!R->qual || !R->qual_off || !R->name_prev))
            return fail(ctx, PSS_EINVAL, "qual / qual_off / name_prev are required for this request");
    }
    return PSS_OK;
}

static cudaEvent_t next_event(pss_ctx *ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

// a pair of timing events tagged `tag` (a pileup statistic id or a PSS_KTAG_* constant), or NULLs when timing is off
static void timing_pair(pss_ctx *ctx, int tag, cudaEvent_t *e0, cudaEvent_t *e1) {
    *e0 = *e1 = nullptr;
    if (!ctx->timing) return;
    cudaEvent_t a = next_event(ctx), b = next_event(ctx);
    if (!a || !b) return;
    size_t pair = ctx->ev_used / 2 - 1;
    if (ctx->ev_stat.size() <= pair) ctx->ev_stat.resize(pair + 1);
    ctx->ev_stat[pair] = tag;
    *e0 = a; *e1 = b;
}

// k_prep + depth cap (+ overlap tweak): everything that depends on the batch and on the
// read-level parameters but not on the output position range.
static int prepare_batch(pss_ctx *ctx, const pss_reads *R, int stepper, int max_depth, int min_mapq, int no_dup,
                         bool need_qual, uint8_t *qual_dst, int32_t *pair_dst, cudaStream_t s, int64_t qual_lo = 0) {
    int64_t n = R->n_reads;
    int rc;
    if ((rc = ensure(ctx, ctx->ri, sizeof(RInfo) * (size_t)n)) != PSS_OK) return rc;
    RInfo *ri = (RInfo *)ctx->ri.p;
    SmallState *st = (SmallState *)ctx->small.p;
    cudaEvent_t p0, p1;
    timing_pair(ctx, PSS_KTAG_PREP, &p0, &p1);
    if (p0) CU(cudaEventRecord(p0, s));
    CU(launch_prep(*R, stepper, min_mapq, no_dup, ri, st, s, &ctx->launches));
    int maxcnt = max_depth ? max_depth : 8000;   // A.0
    CU(launch_admit(ri, R->pos, n, maxcnt, st, s, &ctx->launches));
    if (p1) CU(cudaEventRecord(p1, s));
    ctx->desc_start = ctx->desc_end = -1;   // tile descriptors belong to the previous batch
    ctx->lowq_minq = -1;
    if (need_qual && n > 0) {
        if (!qual_dst) {
            if ((rc = ensure(ctx, ctx->qual, (size_t)R->n_qual_bytes)) != PSS_OK) return rc;
            qual_dst = (uint8_t *)ctx->qual.p - qual_lo;
            ctx->qual_bias = qual_lo;
        }
        if (!pair_dst) {
            if ((rc = ensure(ctx, ctx->pair, sizeof(int32_t) * (size_t)n)) != PSS_OK) return rc;
            pair_dst = (int32_t *)ctx->pair.p;
        }
        if ((rc = ensure(ctx, ctx->next, sizeof(int32_t) * (size_t)n)) != PSS_OK) return rc;
        if (R->n_qual_bytes > 0)
            CU(cudaMemcpyAsync(qual_dst + qual_lo, R->qual + qual_lo, (size_t)R->n_qual_bytes, cudaMemcpyDeviceToDevice, s));
        CU(launch_overlap(*R, ri, qual_dst, pair_dst, (int32_t *)ctx->next.p, s, &ctx->launches));
    }
    return PSS_OK;
}

// the pileup kernel (+ gc) for output positions [start, end); out[f * stride + (p - start)]
static int run_range(pss_ctx *ctx, const pss_reads *R, const pss_pileup_params *P, bool need_qual, int64_t start,
                     int64_t end, int32_t *out, int64_t stride, uint8_t *rowmask, cudaStream_t s, int narrow = 0,
                     int32_t *ovf8 = nullptr) {
    int64_t npos = end - start;
    if (npos <= 0) return PSS_OK;
    // counting statistics without a base-quality filter take the tile (scatter) kernel,
    // everything else the generic per-position kernel
    const bool tile = tile_kernel_supports(P->stat, P->min_baseq, P->no_del) && !(P->flags & (PSS_FLAG_GENERIC_KERNEL | PSS_FLAG_LONG_READS));
    const int tsize = tile ? PSS_TTILE : PSS_TILE;
    int64_t ntiles = (npos + tsize - 1) / tsize;
    if (ntiles > 0x7fffffff) return fail(ctx, PSS_EINVAL, "position range too large for one call");
    int rc;
    PileupArgs a;
    a.ri = (const RInfo *)ctx->ri.p;
    a.cigar = R->cigar;
    a.seq4 = R->seq4;
    a.qual = need_qual ? (const uint8_t *)ctx->qual.p - ctx->qual_bias : nullptr;
    a.qual_off = need_qual ? R->qual_off : nullptr;
    a.seq_off = R->seq_off;
    a.st_ro = (const SmallState *)ctx->small.p;
    a.toolong = &((SmallState *)ctx->small.p)->toolong;
    a.tlen = (P->stat == PSS_TLEN || P->stat == PSS_TLEN_STRAND) ? R->tlen : nullptr;
    a.tile_rng = nullptr;
    a.redo = nullptr;
    a.lowq = nullptr;
    a.n_reads = R->n_reads;
    a.ref = P->ref; a.ref_start = P->ref_start; a.ref_len = P->ref_len;
    a.start = start; a.end = end;
    a.min_baseq = P->min_baseq; a.no_del = P->no_del;
    a.out = out; a.out_stride = stride; a.rowmask = rowmask;
    a.narrow = (tile && tile_kernel_counting(P->stat)) ? narrow : 0;      // 1: uint16 columns, 2: uint8 columns + overflow flag
    a.ovf8 = ovf8;
    cudaEvent_t e0, e1;
    timing_pair(ctx, P->stat, &e0, &e1);
    if (tile) {
        if ((rc = ensure(ctx, ctx->tdesc, tile_desc_bytes(ntiles))) != PSS_OK) return rc;
        if ((rc = ensure(ctx, ctx->refpack, tile_refpack_bytes(ntiles))) != PSS_OK) return rc;
        int need_raw = (P->min_mapq > 0 || P->no_dup || P->no_del || P->min_baseq > 0) ? 1 : 0;
        if (P->min_baseq > 0) {       // mark the bases below the threshold once per batch and threshold
            if (ctx->lowq_minq != P->min_baseq) {
                if ((rc = ensure(ctx, ctx->lowq, tile_lowq_bytes(R->n_reads, R->n_seq_bytes))) != PSS_OK) return rc;
                CU(launch_lowq(*R, a.qual, P->min_baseq, ctx->lowq.p, s, &ctx->launches));
                ctx->lowq_minq = P->min_baseq;
            }
            a.lowq = (const uint32_t *)ctx->lowq.p;
        }
        // descriptors depend on the batch and the position range only: reused by the next statistic
        const bool desc_valid = ctx->desc_start == start && ctx->desc_end == end;
        const bool is_tlen = P->stat == PSS_TLEN || P->stat == PSS_TLEN_STRAND;
        if (is_tlen) {
            if ((rc = ensure(ctx, ctx->redo, tile_redo_bytes(ntiles))) != PSS_OK) return rc;
            CU(cudaMemsetAsync(ctx->redo.p, 0, sizeof(int32_t), s));        // the list's length
            a.redo = (int32_t *)ctx->redo.p;
        }
        CU(launch_tile(P->stat, a, R->seq_off, R->n_reads, (SmallState *)ctx->small.p, need_raw, ntiles, ctx->tdesc.p,
                       ctx->refpack.p, desc_valid, s, &ctx->launches, e0, e1));
        ctx->desc_start = start; ctx->desc_end = end;
        if (is_tlen) {
            // the in-order Welford replay for the tiles whose std_tlen is too close to a rounding boundary to be
            // decided from the exact sums: a fixed small grid walking the list k_tilet wrote
            const int64_t gtiles = (npos + PSS_TILE - 1) / PSS_TILE;
            // candidate reads of a listed 256-position tile: those of the 512-position tile it is half of, straight from
            // the tile kernel's descriptors (a superset in file order is all the replay needs)
            a.tile_rng = (const int2 *)ctx->tdesc.p;
            CU(launch_pileup(P->stat, a, gtiles < 1184 ? gtiles : 1184, true, s, &ctx->launches));     // 148 SMs x 8
            a.redo = nullptr;
        }
    } else {
        if ((rc = ensure(ctx, ctx->rng, sizeof(int2) * (size_t)ntiles)) != PSS_OK) return rc;
        CU(launch_tile_ranges((const RInfo *)ctx->ri.p, R->n_reads, (const SmallState *)ctx->small.p, start, ntiles,
                              tsize, (int2 *)ctx->rng.p, s, &ctx->launches));
        a.tile_rng = (const int2 *)ctx->rng.p;
        if (e0) CU(cudaEventRecord(e0, s));
        CU(launch_pileup(P->stat, a, ntiles, !(P->flags & PSS_FLAG_LONG_READS), s, &ctx->launches));
        if (e1) CU(cudaEventRecord(e1, s));
    }
    if (P->stat == PSS_COVERAGE_GC) {
        cudaEvent_t g0, g1;
        timing_pair(ctx, PSS_KTAG_GC, &g0, &g1);
        if (g0) CU(cudaEventRecord(g0, s));
        CU(launch_gc(P->ref, P->ref_start, P->ref_len, P->contig_len, start, end, P->window_size, P->window_offset,
                     out, s, &ctx->launches));
        if (g1) CU(cudaEventRecord(g1, s));
    }
    return PSS_OK;
}

static int check_params(pss_ctx *ctx, const pss_pileup_params *P) {
    if (!P) return fail(ctx, PSS_EINVAL, "params is NULL");
    if (P->stat < 0 || P->stat >= PSS_N_PILEUP_STATS) return fail(ctx, PSS_EINVAL, "unsupported statistics type: %d", P->stat);
    if (P->stepper < 0 || P->stepper > 2) return fail(ctx, PSS_EINVAL, "unknown stepper: %d", P->stepper);
    if (P->end < P->start) return fail(ctx, PSS_EINVAL, "end < start");
    if (P->max_depth < 0) return fail(ctx, PSS_EINVAL, "max_depth < 0");
    if (pss_pileup_needs_ref(P->stat) && (!P->ref && P->ref_len > 0)) return fail(ctx, PSS_EINVAL, "ref is NULL");
    if (P->stat == PSS_COVERAGE_GC && P->window_size <= 0) return fail(ctx, PSS_EINVAL, "window_size <= 0");
    return PSS_OK;
}

extern "C" int pss_prepare(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth, int32_t min_mapq,
                           int32_t no_dup, int32_t with_qual, void *stream) {
    if (!ctx) return PSS_EINVAL;
    ctx->prepared = false;
    int rc;
    if (stepper < 0 || stepper > 2) return fail(ctx, PSS_EINVAL, "unknown stepper: %d", stepper);
    if (max_depth < 0) return fail(ctx, PSS_EINVAL, "max_depth < 0");
    if ((rc = check_reads(ctx, reads, with_qual != 0)) != PSS_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(ctx->device));
    if ((rc = prepare_batch(ctx, reads, stepper, max_depth, min_mapq, no_dup, with_qual != 0, nullptr, nullptr, s)) != PSS_OK)
        return rc;
    ctx->prep_reads = *reads;
    ctx->prep_stepper = stepper; ctx->prep_max_depth = max_depth; ctx->prep_min_mapq = min_mapq;
    ctx->prep_no_dup = no_dup ? 1 : 0; ctx->prep_qual = with_qual ? 1 : 0;
    ctx->prepared = true;
    return PSS_OK;
}

extern "C" int pss_pileup_prepared(pss_ctx *ctx, const pss_pileup_params *P, int32_t *out, uint8_t *rowmask, void *stream) {
    if (!ctx) return PSS_EINVAL;
    int rc;
    if ((rc = check_params(ctx, P)) != PSS_OK) return rc;
    if (!ctx->prepared) return fail(ctx, PSS_EINVAL, "no prepared batch: call pss_prepare first");
    if (P->stepper != ctx->prep_stepper || P->max_depth != ctx->prep_max_depth || P->min_mapq != ctx->prep_min_mapq ||
        (P->no_dup ? 1 : 0) != ctx->prep_no_dup)
        return fail(ctx, PSS_EINVAL, "stepper / max_depth / min_mapq / no_dup differ from the prepared batch");
    bool need_qual = pss_pileup_needs_qual(P->stat, P->min_baseq);
    if (need_qual && !ctx->prep_qual) return fail(ctx, PSS_EINVAL, "the batch was prepared without base qualities");
    if (P->end > P->start && (!out || !rowmask)) return fail(ctx, PSS_EINVAL, "out / rowmask is NULL");
    CU(cudaSetDevice(ctx->device));
    return run_range(ctx, &ctx->prep_reads, P, need_qual, P->start, P->end, out, P->end - P->start, rowmask,
                     (cudaStream_t)stream);
}

extern "C" int pss_pileup(pss_ctx *ctx, const pss_reads *reads, const pss_pileup_params *P, int32_t *out,
                          uint8_t *rowmask, void *stream) {
    if (!ctx) return PSS_EINVAL;
    int rc;
    if ((rc = check_params(ctx, P)) != PSS_OK) return rc;
    bool need_qual = pss_pileup_needs_qual(P->stat, P->min_baseq);
    if ((rc = pss_prepare(ctx, reads, P->stepper, P->max_depth, P->min_mapq, P->no_dup, need_qual, stream)) != PSS_OK)
        return rc;
    return pss_pileup_prepared(ctx, P, out, rowmask, stream);
}

extern "C" int pss_admit(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth, uint8_t *admitted,
                         void *stream) {
    if (!ctx) return PSS_EINVAL;
    int rc;
    if ((rc = check_reads(ctx, reads, false)) != PSS_OK) return rc;
    if (stepper < 0 || stepper > 2) return fail(ctx, PSS_EINVAL, "unknown stepper: %d", stepper);
    if (reads->n_reads > 0 && !admitted) return fail(ctx, PSS_EINVAL, "admitted is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    if ((rc = prepare_batch(ctx, reads, stepper, max_depth, 0, 0, false, nullptr, nullptr, s)) != PSS_OK) return rc;
    CU(launch_export_admitted((const RInfo *)ctx->ri.p, reads->n_reads, admitted, s, &ctx->launches));
    return PSS_OK;
}

extern "C" int pss_overlap_tweak(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth,
                                 uint8_t *qual_out, int32_t *pair_first, void *stream) {
    if (!ctx) return PSS_EINVAL;
    int rc;
    if ((rc = check_reads(ctx, reads, true)) != PSS_OK) return rc;
    if (stepper < 0 || stepper > 2) return fail(ctx, PSS_EINVAL, "unknown stepper: %d", stepper);
    if (reads->n_reads > 0 && !qual_out) return fail(ctx, PSS_EINVAL, "qual_out is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    return prepare_batch(ctx, reads, stepper, max_depth, 0, 0, true, qual_out, pair_first, s);
}

// ---------------------------------------------------------------------- pileup (host)

template <class T>
static int to_device(pss_ctx *ctx, DevBuf &b, const T *host, size_t count, const T **dev, cudaStream_t s) {
    *dev = nullptr;
    if (!host) return PSS_OK;
    int rc = ensure(ctx, b, sizeof(T) * count);
    if (rc != PSS_OK) return rc;
    if (count) {
        CU(cudaMemcpyAsync(b.p, host, sizeof(T) * count, cudaMemcpyHostToDevice, s));
        ctx->h2d_bytes += (int64_t)(sizeof(T) * count);
    }
    *dev = (const T *)b.p;
    return PSS_OK;
}

// Ragged payload (cigar / seq4 / qual): only the span [off[0], off[n]) the batch refers to is
// copied, and the device pointer is biased so the (absolute) offsets keep working.  This is
// what lets a caller pass a zero-copy sub-range view of a larger host batch.
template <class T>
static int to_device_span(pss_ctx *ctx, DevBuf &b, const T *host, const int64_t *off, size_t n, const T **dev,
                          cudaStream_t s) {
    *dev = nullptr;
    if (!host || !off) return PSS_OK;
    int64_t lo = off[0], hi = off[n];
    if (hi < lo) return fail(ctx, PSS_EINVAL, "offsets decrease");
    const T *d;
    int rc = to_device(ctx, b, host + lo, (size_t)(hi - lo), &d, s);
    if (rc != PSS_OK) return rc;
    *dev = d - lo;
    return PSS_OK;
}

static int upload_reads(pss_ctx *ctx, const pss_reads *H, pss_reads *D, bool need_seq, bool need_qual, cudaStream_t s,
                        DevBuf *hb = nullptr) {
    if (!hb) hb = ctx->hb;
    *D = *H;
    size_t n = (size_t)H->n_reads;
    int rc;
#define UP(idx, field, count) if ((rc = to_device(ctx, hb[idx], H->field, (count), &D->field, s)) != PSS_OK) return rc
#define UPS(idx, field, off) if ((rc = to_device_span(ctx, hb[idx], H->field, H->off, n, &D->field, s)) != PSS_OK) return rc
    UP(0, pos, n); UP(1, flag, n); UP(2, mapq, n); UP(3, tlen, n); UP(4, mtid, n); UP(5, l_seq, n);
    UP(6, cigar_off, n + 1); UP(8, seq_off, n + 1);
    if (n > 0) { UPS(7, cigar, cigar_off); }
    if (need_seq && n > 0) { UPS(9, seq4, seq_off); }
    else {   // never dereferenced, but must be non-NULL for the argument check
        if ((rc = ensure(ctx, hb[9], 16)) != PSS_OK) return rc;
        D->seq4 = (const uint8_t *)hb[9].p;
    }
    if (need_qual && n > 0) {
        UP(10, qual_off, n + 1); UPS(11, qual, qual_off); UP(12, name_prev, n);
        D->n_qual_bytes = H->qual_off[n] - H->qual_off[0];
    }
    else { D->qual_off = nullptr; D->qual = nullptr; D->name_prev = nullptr; }
#undef UP
#undef UPS
    return PSS_OK;
}

static bool stat_reads_seq(int stat) {
    return stat == PSS_VARIATION || stat == PSS_VARIATION_STRAND || stat == PSS_BASEQ_EXT || stat == PSS_BASEQ_EXT_STRAND;
}

// One statistic over [start, end), int32 columns straight to the caller's host buffer: position
// chunks of <= ~256 MB of output, two device buffers alternating against the D2H copies.
static int host_stat_wide(pss_ctx *ctx, const pss_reads *D, const pss_pileup_params *Pd, bool q, int32_t *out,
                          int64_t ostride, uint8_t *rowmask, cudaStream_t s, int *slot) {
    int rc;
    const int nf = kPileupFields[Pd->stat];
    const int64_t npos = Pd->end - Pd->start;
    int64_t chunk = ((int64_t)(256ll << 20) / (4ll * nf)) / PSS_TTILE * PSS_TTILE;
    if (chunk > npos) chunk = (npos + PSS_TTILE - 1) / PSS_TTILE * PSS_TTILE;
    if (chunk < PSS_TTILE) chunk = PSS_TTILE;
    int &k = *slot;
    for (int64_t cs = Pd->start; cs < Pd->end; cs += chunk, ++k) {
        int64_t ce = cs + chunk < Pd->end ? cs + chunk : Pd->end;
        int64_t cn = ce - cs;
        int b = k & 1;
        const int64_t stride = (chunk + 3) & ~(int64_t)3;   // padded rows: 128-bit stores everywhere
        if (k >= 2) CU(cudaStreamWaitEvent(s, ctx->ev_copied[b], 0));   // buffer free again
        if (ctx->outbuf[b].bytes < sizeof(int32_t) * (size_t)nf * (size_t)stride || ctx->maskbuf[b].bytes < (size_t)stride) {
            // growing a buffer frees it: wait for the copy that may still read it
            CU(cudaStreamSynchronize(ctx->copy_stream));
            if ((rc = ensure(ctx, ctx->outbuf[b], sizeof(int32_t) * (size_t)nf * (size_t)stride)) != PSS_OK) return rc;
            if ((rc = ensure(ctx, ctx->maskbuf[b], (size_t)stride)) != PSS_OK) return rc;
        }
        if ((rc = run_range(ctx, D, Pd, q, cs, ce, (int32_t *)ctx->outbuf[b].p, stride, (uint8_t *)ctx->maskbuf[b].p, s)) != PSS_OK)
            return rc;
        CU(cudaEventRecord(ctx->ev_done[b], s));
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
        CU(cudaMemcpy2DAsync(out + (cs - Pd->start), sizeof(int32_t) * (size_t)ostride, ctx->outbuf[b].p,
                             sizeof(int32_t) * (size_t)stride, sizeof(int32_t) * (size_t)cn, (size_t)nf,
                             cudaMemcpyDeviceToHost, ctx->copy_stream));
        CU(cudaMemcpyAsync(rowmask + (cs - Pd->start), ctx->maskbuf[b].p, (size_t)cn, cudaMemcpyDeviceToHost,
                           ctx->copy_stream));
        CU(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
        ctx->d2h_bytes += (int64_t)(sizeof(int32_t) * (size_t)cn * (size_t)nf + (size_t)cn);
    }
    return PSS_OK;
}

#define PSS_ERETRY_WIDE (-100)   // internal: a tile needed several passes, narrow transport cannot express it

// Same, for the statistics of the tile kernel.  What bounds the host path is host memory
// traffic (measured on the GPU box: ~160 GB/s over 16 cores, scripts/host_probe.cu): every
// result byte is written by the DMA engine into a pinned staging buffer, read back by a worker
// and written as int32 into the caller's column.  So the device writes the NARROWEST columns
// that hold the counts: 8-bit when every count of a position chunk is < 256 (any ordinary WGS
// depth), else the chunk is redone with 16-bit columns (< 65536: every single-pass tile), else
// the caller repeats the statistic on the int32 pipeline.  A quarter / half of the bytes cross
// PCIe, and the worker pool widens them into the caller's int32 columns while the next chunks
// are computed and copied.  Three buffers rotate: compute k, copy k - 1, widen k - 2.
struct RowSink {              // records instead of columns (pss_pileup_host_rows)
    pss_row_pack pack;
    int64_t pad_start = 0, pad_end = 0;
    int64_t capacity = 0, n_rows = 0;
};

struct NarrowStat {           // one statistic of a host call on the narrow transport
    pss_pileup_params Pd;     // with the device reference pointer
    bool q;
    int32_t *out;
    uint8_t *rowmask;
    int esz_fixed = 0;        // 4: int32 columns through the same pipeline (statistics without narrow output; sink only)
    RowSink *sink = nullptr;  // chunks are packed into records instead of widened into `out`
};

// All the narrow statistics of a call share ONE pipeline: the position chunks of statistic i + 1 follow those of
// statistic i through the three buffers without a drain in between (a drain costs the compute + copy of a first
// chunk and the widening of a last one, per statistic and call).
static int host_stats_narrow(pss_ctx *ctx, const pss_reads *D, const std::vector<NarrowStat> &stats, int64_t ostride,
                             cudaStream_t s) {
    int rc;
    if (stats.empty()) return PSS_OK;
    struct Item { int st; int64_t cs, cn, stride; };
    std::vector<Item> items;
    size_t max_bytes16 = 0;
    int64_t max_stride = 0;
    for (size_t i = 0; i < stats.size(); ++i) {
        const pss_pileup_params &Pd = stats[i].Pd;
        const int nf = kPileupFields[Pd.stat];
        const int64_t npos = Pd.end - Pd.start;
        const int64_t bpe = stats[i].esz_fixed == 4 ? 4 : 2;
        int64_t chunk = ((int64_t)(128ll << 20) / (bpe * nf)) / PSS_TTILE * PSS_TTILE;
        // a short range still gets ~8 chunks (>= 1 Mi positions each) so that compute, copy and host work overlap
        const int64_t eighth = std::max<int64_t>(1 << 20, ((npos + 7) / 8 + PSS_TTILE - 1) / PSS_TTILE * PSS_TTILE);
        if (chunk > eighth) chunk = eighth;
        if (const char *hc = getenv("PSS_HOST_CHUNK"))      // positions per chunk (tests: many chunks on small inputs)
            if (atoll(hc) > 0) chunk = (atoll(hc) + PSS_TTILE - 1) / PSS_TTILE * PSS_TTILE;
        if (chunk > npos) chunk = (npos + PSS_TTILE - 1) / PSS_TTILE * PSS_TTILE;
        if (chunk < PSS_TTILE) chunk = PSS_TTILE;
        for (int64_t cs = Pd.start; cs < Pd.end; cs += chunk)
            items.push_back({(int)i, cs, (cs + chunk < Pd.end ? cs + chunk : Pd.end) - cs, chunk});   // stride = chunk: rows 8-byte aligned
        max_bytes16 = std::max(max_bytes16, (size_t)bpe * (size_t)nf * (size_t)chunk);
        max_stride = std::max(max_stride, chunk);
    }
    static const bool no_narrow8 = getenv("PSS_NO_NARROW8") != nullptr;
    if (!ctx->pool) ctx->pool = pss_pool_create(0);
    CU(cudaStreamSynchronize(ctx->copy_stream));        // nothing of an earlier call is in flight
    for (int b = 0; b < 3; ++b) {
        if ((rc = ensure(ctx, ctx->nbuf[b], max_bytes16)) != PSS_OK) return rc;
        if ((rc = ensure(ctx, ctx->nmask[b], (size_t)max_stride)) != PSS_OK) return rc;
        if (ctx->hstage_bytes[b] < max_bytes16) {
            if (ctx->hstage[b]) cudaFreeHost(ctx->hstage[b]);
            ctx->hstage[b] = nullptr; ctx->hstage_bytes[b] = 0;
            CU(cudaMallocHost(&ctx->hstage[b], max_bytes16));
            ctx->hstage_bytes[b] = max_bytes16;
        }
        if (!ctx->nev_done[b]) {
            CU(cudaEventCreateWithFlags(&ctx->nev_done[b], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->nev_copied[b], cudaEventDisableTiming));
        }
    }
    if ((rc = ensure(ctx, ctx->nflag, 4 * sizeof(int32_t))) != PSS_OK) return rc;
    if (!ctx->hflag) CU(cudaMallocHost((void **)&ctx->hflag, 4 * sizeof(int32_t)));
    int32_t *dflag = (int32_t *)ctx->nflag.p;
    // clear the overflow flag of this batch's device state
    CU(cudaMemsetAsync(&((SmallState *)ctx->small.p)->overflow16, 0, sizeof(int32_t), s));
    size_t ticket[3] = {0, 0, 0};
    int held[3] = {-1, -1, -1};      // item in buffer b
    int esz[3] = {2, 2, 2};
    int cool = 0;          // chunks left that go straight to 16 bits after an 8-bit overflow
    pss_pool *pool = ctx->pool;
    // compute item it into device buffer b with `e`-byte elements and copy it to staging buffer b
    auto produce = [&](int b, int it, int e) -> int {
        const Item &I = items[(size_t)it];
        const NarrowStat &S = stats[(size_t)I.st];
        const int nf = kPileupFields[S.Pd.stat];
        if (S.esz_fixed) e = S.esz_fixed;
        if (e == 1) CU(cudaMemsetAsync(dflag + b, 0, sizeof(int32_t), s));
        int r = run_range(ctx, D, &S.Pd, S.q, I.cs, I.cs + I.cn, (int32_t *)ctx->nbuf[b].p, I.stride, (uint8_t *)ctx->nmask[b].p, s,
                          e == 1 ? 2 : (e == 2 ? 1 : 0), e == 1 ? dflag + b : nullptr);
        if (r != PSS_OK) return r;
        const size_t bytes = (size_t)e * (size_t)nf * (size_t)I.stride;
        CU(cudaEventRecord(ctx->nev_done[b], s));
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->nev_done[b], 0));
        CU(cudaMemcpyAsync(ctx->hstage[b], ctx->nbuf[b].p, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CU(cudaMemcpyAsync(S.rowmask + (I.cs - S.Pd.start), ctx->nmask[b].p, (size_t)I.cn, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (e == 1) CU(cudaMemcpyAsync(ctx->hflag + b, dflag + b, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
        CU(cudaEventRecord(ctx->nev_copied[b], ctx->copy_stream));
        ctx->d2h_bytes += (int64_t)(bytes + (size_t)I.cn);
        held[b] = it; esz[b] = e;
        return PSS_OK;
    };
    // the D2H of buffer b has been enqueued: wait for it, redo the chunk with 16 bits if its counts
    // did not fit 8, then hand its rows to the pool
    auto consume = [&](int b) -> int {
        CU(cudaEventSynchronize(ctx->nev_copied[b]));
        if (esz[b] == 1 && ctx->hflag[b]) {
            cool = 8;
            int r = produce(b, held[b], 2);
            if (r != PSS_OK) return r;
            CU(cudaEventSynchronize(ctx->nev_copied[b]));
        }
        const Item &I = items[(size_t)held[b]];
        const NarrowStat &S = stats[(size_t)I.st];
        const int nf = kPileupFields[S.Pd.stat];
        const uint8_t *src = (const uint8_t *)ctx->hstage[b];
        const int e = esz[b];
        size_t t = 0;
        if (S.sink) {
            // rows: count the emitted positions of every piece (this thread, in position order), then the
            // workers pack the pieces into the records from their first row on
            RowSink *K = S.sink;
            uint8_t *m = S.rowmask + (I.cs - S.Pd.start);
            const int64_t pa = std::max(K->pad_start, I.cs), pz = std::min(K->pad_end, I.cs + I.cn);
            if (pz > pa) memset(m + (pa - I.cs), 1, (size_t)(pz - pa));
            const int64_t rpiece = 1 << 15;
            for (int64_t o = 0; o < I.cn; o += rpiece) {
                const int64_t len = std::min(rpiece, I.cn - o);
                const int64_t cnt = pss_count_rows(m + o, len);
                if (cnt == 0) continue;
                if (K->n_rows + cnt > K->capacity) return fail(ctx, PSS_EINVAL, "rows: more than capacity_rows (%lld) rows", (long long)K->capacity);
                const int64_t row0 = K->n_rows;
                K->n_rows += cnt;
                const uint8_t *sp = src + (size_t)o * e, *mp = m + o;
                const int64_t stride = I.stride, p0 = I.cs + o;
                pss_row_pack *pk = &K->pack;
                t = pss_pool_submit(pool, [pk, sp, stride, e, mp, p0, len, row0] { pss_pack_piece(pk, sp, stride, e, mp, p0, len, row0); });
            }
            ticket[b] = t;
            return PSS_OK;
        }
        const int64_t piece = 1 << 20;
        for (int f = 0; f < nf; ++f)
            for (int64_t o = 0; o < I.cn; o += piece) {
                const uint8_t *sp = src + ((size_t)f * I.stride + o) * e;
                int32_t *dp = S.out + (size_t)f * (size_t)ostride + (I.cs - S.Pd.start) + o;
                size_t len = (size_t)((I.cn - o) < piece ? (I.cn - o) : piece);
                if (e == 1) t = pss_pool_submit(pool, [sp, dp, len] { pss_widen_u8(sp, dp, len); });
                else t = pss_pool_submit(pool, [sp, dp, len] { pss_widen_u16((const uint16_t *)sp, dp, len); });
            }
        ticket[b] = t;
        return PSS_OK;
    };
    const int n_items = (int)items.size();
    for (int k = 0; k < n_items; ++k) {
        const int b = k % 3;
        if (k >= 3) {
            pss_pool_wait(pool, ticket[b]);                          // staging buffer b widened
            CU(cudaStreamWaitEvent(s, ctx->nev_copied[b], 0));        // device buffer b copied out
        }
        int e = 1;
        if (no_narrow8) e = 2;
        else if (cool > 0) { e = 2; --cool; }
        if ((rc = produce(b, k, e)) != PSS_OK) return rc;
        if (k >= 1 && (rc = consume((k - 1) % 3)) != PSS_OK) return rc;   // widen the previous chunk meanwhile
    }
    if (n_items >= 1 && (rc = consume((n_items - 1) % 3)) != PSS_OK) return rc;
    for (int b = 0; b < 3; ++b) pss_pool_wait(pool, ticket[b]);
    // a tile that needed several passes (> 65535 reads) cannot be expressed in 16 bits
    int32_t ovf = 0;
    CU(cudaMemcpyAsync(&ovf, &((SmallState *)ctx->small.p)->overflow16, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return ovf ? PSS_ERETRY_WIDE : PSS_OK;
}

// One batch (already the reads of [P[0].start, P[0].end) plus their halo), all statistics: prepare once, then
// every statistic through the narrow / int32 result pipelines.  D: device copy of the batch, on stream s.
// outs[i] rows are `ostride` elements apart.
static int host_multi_run(pss_ctx *ctx, const pss_reads *H, const pss_reads *D, int32_t n_stats, const pss_pileup_params *P,
                          const uint8_t *ref_dev, bool need_qual, int32_t *const *outs, int64_t ostride, uint8_t *const *rowmasks,
                          cudaStream_t s) {
    int rc;
    if ((rc = prepare_batch(ctx, D, P[0].stepper, P[0].max_depth, P[0].min_mapq, P[0].no_dup, need_qual, nullptr, nullptr,
                            s, (need_qual && H->n_reads > 0) ? H->qual_off[0] : 0)) != PSS_OK)
        return rc;
    int k = 0;     // running buffer slot of the int32 pipeline
    static const bool no_narrow = getenv("PSS_NO_NARROW") != nullptr;
    // the counting statistics of the tile kernel travel narrow, all of them through one pipeline
    // (coverage_gc's gc column is written as int32 by k_gc, so it stays on the int32 pipeline)
    std::vector<NarrowStat> narrow;
    std::vector<char> is_narrow((size_t)n_stats, 0);
    for (int i = 0; i < n_stats; ++i) {
        if (no_narrow || !tile_kernel_counting(P[i].stat) || P[i].stat == PSS_COVERAGE_GC ||
            (P[i].flags & (PSS_FLAG_GENERIC_KERNEL | PSS_FLAG_LONG_READS)))
            continue;
        NarrowStat ns;
        ns.Pd = P[i];
        ns.Pd.ref = (P[i].ref && P[i].ref_len > 0) ? ref_dev : nullptr;
        ns.q = pss_pileup_needs_qual(P[i].stat, P[i].min_baseq);
        ns.out = outs[i]; ns.rowmask = rowmasks[i];
        narrow.push_back(ns);
        is_narrow[(size_t)i] = 1;
    }
    rc = host_stats_narrow(ctx, D, narrow, ostride, s);
    if (rc == PSS_ERETRY_WIDE) std::fill(is_narrow.begin(), is_narrow.end(), 0);      // > 65535 reads in a tile: int32 for all
    else if (rc != PSS_OK) return rc;
    for (int i = 0; i < n_stats; ++i) {
        if (is_narrow[(size_t)i]) continue;
        pss_pileup_params Pd = P[i];
        Pd.ref = (P[i].ref && P[i].ref_len > 0) ? ref_dev : nullptr;
        const bool q = pss_pileup_needs_qual(P[i].stat, P[i].min_baseq);
        if ((rc = host_stat_wide(ctx, D, &Pd, q, outs[i], ostride, rowmasks[i], s, &k)) != PSS_OK) return rc;
    }
    CU(cudaStreamSynchronize(ctx->copy_stream));
    return PSS_OK;
}

static int check_multi(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats, const pss_pileup_params *P,
                       int32_t *const *outs, uint8_t *const *rowmasks, bool *need_qual, bool *need_seq) {
    if (n_stats < 1 || !P || !outs || !rowmasks) return fail(ctx, PSS_EINVAL, "n_stats < 1 or a NULL array");
    int rc;
    *need_qual = *need_seq = false;
    for (int i = 0; i < n_stats; ++i) {
        if ((rc = check_params(ctx, &P[i])) != PSS_OK) return rc;
        if (P[i].stepper != P[0].stepper || P[i].max_depth != P[0].max_depth || P[i].min_mapq != P[0].min_mapq ||
            (P[i].no_dup != 0) != (P[0].no_dup != 0) || P[i].start != P[0].start || P[i].end != P[0].end)
            return fail(ctx, PSS_EINVAL, "statistics of one call must share stepper, max_depth, min_mapq, no_dup, start, end");
        if (P[i].ref && P[0].ref && (P[i].ref != P[0].ref || P[i].ref_start != P[0].ref_start || P[i].ref_len != P[0].ref_len))
            return fail(ctx, PSS_EINVAL, "statistics of one call must share the reference span");
        bool q = pss_pileup_needs_qual(P[i].stat, P[i].min_baseq);
        *need_qual |= q;
        *need_seq |= q || stat_reads_seq(P[i].stat);   // the overlap tweak compares the mates' bases
        if (P[i].end > P[i].start && (!outs[i] || !rowmasks[i])) return fail(ctx, PSS_EINVAL, "out / rowmask is NULL");
    }
    return check_reads(ctx, reads, *need_qual);
}

extern "C" int pss_pileup_host_multi(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats, const pss_pileup_params *P,
                                     int32_t *const *outs, uint8_t *const *rowmasks) {
    if (!ctx) return PSS_EINVAL;
    int rc;
    bool need_qual = false, need_seq = false;
    if ((rc = check_multi(ctx, reads, n_stats, P, outs, rowmasks, &need_qual, &need_seq)) != PSS_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    cudaStream_t s = nullptr;   // default stream for compute, ctx->copy_stream for D2H
    pss_reads D;
    if ((rc = upload_reads(ctx, reads, &D, need_seq, need_qual, s)) != PSS_OK) return rc;
    const uint8_t *ref_dev = nullptr;
    for (int i = 0; i < n_stats && !ref_dev; ++i)
        if (P[i].ref && P[i].ref_len > 0)
            if ((rc = to_device(ctx, ctx->hb[13], P[i].ref, (size_t)P[i].ref_len, &ref_dev, s)) != PSS_OK) return rc;
    if ((rc = host_multi_run(ctx, reads, &D, n_stats, P, ref_dev, need_qual, outs, P[0].end - P[0].start, rowmasks, s)) != PSS_OK)
        return rc;
    rc = pss_synchronize(ctx, s);
    if (rc == PSS_ETOOLONG) {
        // a read did not fit the kernels' staging buffers (long-read data): repeat the call on the per-position
        // kernel that reads bases / qualities from global memory -- callers of the reference API cannot set flags
        bool all_long = true;
        std::vector<pss_pileup_params> Pl(P, P + n_stats);
        for (auto &q : Pl) { all_long = all_long && (q.flags & PSS_FLAG_LONG_READS); q.flags |= PSS_FLAG_LONG_READS; }
        if (all_long) return rc;
        if ((rc = host_multi_run(ctx, reads, &D, n_stats, Pl.data(), ref_dev, need_qual, outs, P[0].end - P[0].start, rowmasks, s)) != PSS_OK)
            return rc;
        rc = pss_synchronize(ctx, s);
    }
    return rc;
}

// The pairs of the mate-overlap rewrite for a whole batch from its read metadata alone (no sequences, no qualities):
// what lets a caller cut a request that reads base qualities into position pieces -- see include/pss_b200.h.
extern "C" int pss_pairs_host(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth, int32_t *pair_first) {
    if (!ctx) return PSS_EINVAL;
    int rc;
    if ((rc = check_reads(ctx, reads, false)) != PSS_OK) return rc;
    if (stepper < 0 || stepper > 2) return fail(ctx, PSS_EINVAL, "unknown stepper: %d", stepper);
    if (max_depth < 0) return fail(ctx, PSS_EINVAL, "max_depth < 0");
    const int64_t n = reads->n_reads;
    if (n == 0) return PSS_OK;
    if (!reads->name_prev || !pair_first) return fail(ctx, PSS_EINVAL, "name_prev / pair_first is NULL");
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    cudaStream_t s = nullptr;
    pss_reads D;
    if ((rc = upload_reads(ctx, reads, &D, false, false, s)) != PSS_OK) return rc;
    if ((rc = to_device(ctx, ctx->hb[12], reads->name_prev, (size_t)n, &D.name_prev, s)) != PSS_OK) return rc;
    if ((rc = prepare_batch(ctx, &D, stepper, max_depth, 0, 0, false, nullptr, nullptr, s)) != PSS_OK) return rc;
    if ((rc = ensure(ctx, ctx->pair, sizeof(int32_t) * (size_t)n)) != PSS_OK) return rc;
    if ((rc = ensure(ctx, ctx->next, sizeof(int32_t) * (size_t)n)) != PSS_OK) return rc;
    CU(launch_overlap(D, (const RInfo *)ctx->ri.p, nullptr, (int32_t *)ctx->pair.p, (int32_t *)ctx->next.p, s, &ctx->launches));
    CU(cudaMemcpyAsync(pair_first, ctx->pair.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
    ctx->d2h_bytes += (int64_t)(sizeof(int32_t) * (size_t)n);
    return pss_synchronize(ctx, s);
}

// ---------------------------------------------------------------------- pileup rows (host)

// One statistic as packed records: the narrow pipeline with a row sink; statistics without narrow output (the
// quality kinds, tlen, coverage_gc, the generic kernel) and batches with a > 65535-read tile go through the
// same three buffers as int32 columns.
static int host_rows_run(pss_ctx *ctx, const pss_reads *H, const pss_reads *D, const pss_pileup_params *P,
                         const uint8_t *ref_dev, bool need_qual, RowSink *K, uint8_t *mask, cudaStream_t s) {
    int rc;
    if ((rc = prepare_batch(ctx, D, P->stepper, P->max_depth, P->min_mapq, P->no_dup, need_qual, nullptr, nullptr, s,
                            (need_qual && H->n_reads > 0) ? H->qual_off[0] : 0)) != PSS_OK)
        return rc;
    static const bool no_narrow = getenv("PSS_NO_NARROW") != nullptr;
    NarrowStat ns;
    ns.Pd = *P;
    ns.Pd.ref = (P->ref && P->ref_len > 0) ? ref_dev : nullptr;
    ns.q = need_qual;
    ns.out = nullptr; ns.rowmask = mask; ns.sink = K;
    const bool narrow = !no_narrow && tile_kernel_counting(P->stat) && P->stat != PSS_COVERAGE_GC &&
                        !(P->flags & (PSS_FLAG_GENERIC_KERNEL | PSS_FLAG_LONG_READS));
    ns.esz_fixed = narrow ? 0 : 4;
    for (;;) {
        K->n_rows = 0;
        K->pack.n_negative = 0;
        rc = host_stats_narrow(ctx, D, std::vector<NarrowStat>(1, ns), 0, s);
        if (rc == PSS_ERETRY_WIDE && ns.esz_fixed == 0) { ns.esz_fixed = 4; continue; }
        break;
    }
    // no worker still reads the staging buffers / the mask when this returns, whatever happened
    if (ctx->pool) pss_pool_wait(ctx->pool, pss_pool_submit(ctx->pool, [] {}));
    if (rc != PSS_OK) return rc;
    CU(cudaStreamSynchronize(ctx->copy_stream));
    return PSS_OK;
}

extern "C" int pss_pileup_host_rows(pss_ctx *ctx, const pss_reads *reads, const pss_pileup_params *P, const pss_row_layout *L,
                                    void *rows, int64_t capacity_rows, int64_t *n_rows, int64_t *n_negative) {
    if (!ctx) return PSS_EINVAL;
    if (!P || !L || !n_rows) return fail(ctx, PSS_EINVAL, "params / layout / n_rows is NULL");
    int rc;
    *n_rows = 0;
    if (n_negative) *n_negative = 0;
    if ((rc = check_params(ctx, P)) != PSS_OK) return rc;
    const int nf = kPileupFields[P->stat];
    if (L->n_fields != nf || nf > 96) return fail(ctx, PSS_EINVAL, "layout: n_fields %d, the statistic has %d", L->n_fields, nf);
    if (L->row_bytes < 1 || !L->field_off || !L->field_bytes || capacity_rows < 0 || (capacity_rows > 0 && !rows))
        return fail(ctx, PSS_EINVAL, "layout: bad row_bytes / field arrays / rows");
    RowSink K;
    pss_row_pack &pk = K.pack;
    pk.rows = (uint8_t *)rows; pk.row_bytes = L->row_bytes;
    for (int j = 0; j < nf; ++j) {
        if (L->field_off[j] < 0) continue;
        const int fb = L->field_bytes[j];
        if ((fb != 4 && fb != 1) || L->field_off[j] + fb > L->row_bytes) return fail(ctx, PSS_EINVAL, "layout: bad field %d", j);
        pk.f_idx[pk.n_want] = j; pk.f_off[pk.n_want] = L->field_off[j]; pk.f_bytes[pk.n_want] = fb;
        ++pk.n_want;
    }
    if (L->chrom_off >= 0 && (!L->chrom || L->chrom_bytes < 0 || L->chrom_off + L->chrom_bytes > L->row_bytes))
        return fail(ctx, PSS_EINVAL, "layout: bad chrom");
    if ((L->pos_off >= 0 && L->pos_off + 4 > L->row_bytes) || (L->ref_off >= 0 && L->ref_off + 1 > L->row_bytes) || L->check_field >= nf)
        return fail(ctx, PSS_EINVAL, "layout: bad pos / ref offset or check_field");
    pk.chrom = L->chrom; pk.chrom_off = L->chrom_off; pk.chrom_bytes = L->chrom_bytes;
    pk.pos_off = L->pos_off; pk.ref_off = L->ref_off; pk.check_field = L->check_field; pk.pos_add = L->pos_add;
    pk.ref = P->ref; pk.ref_start = P->ref_start; pk.ref_len = P->ref ? P->ref_len : 0;
    K.pad_start = L->pad_start; K.pad_end = L->pad_end; K.capacity = capacity_rows;
    const bool need_qual = pss_pileup_needs_qual(P->stat, P->min_baseq);
    const bool need_seq = need_qual || stat_reads_seq(P->stat);
    if ((rc = check_reads(ctx, reads, need_qual)) != PSS_OK) return rc;
    const int64_t npos = P->end - P->start;
    if (npos <= 0) return PSS_OK;
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    cudaStream_t s = nullptr;
    pss_reads D;
    if ((rc = upload_reads(ctx, reads, &D, need_seq, need_qual, s)) != PSS_OK) return rc;
    const uint8_t *ref_dev = nullptr;
    if (P->ref && P->ref_len > 0)
        if ((rc = to_device(ctx, ctx->hb[13], P->ref, (size_t)P->ref_len, &ref_dev, s)) != PSS_OK) return rc;
    std::vector<uint8_t> mask((size_t)npos);
    if ((rc = host_rows_run(ctx, reads, &D, P, ref_dev, need_qual, &K, mask.data(), s)) != PSS_OK) return rc;
    rc = pss_synchronize(ctx, s);
    if (rc == PSS_ETOOLONG && !(P->flags & PSS_FLAG_LONG_READS)) {     // long reads: the per-position kernel (see pss_pileup_host_multi)
        pss_pileup_params Pl = *P;
        Pl.flags |= PSS_FLAG_LONG_READS;
        if ((rc = host_rows_run(ctx, reads, &D, &Pl, ref_dev, need_qual, &K, mask.data(), s)) != PSS_OK) return rc;
        rc = pss_synchronize(ctx, s);
    }
    if (rc != PSS_OK) return rc;
    *n_rows = K.n_rows;
    if (n_negative) *n_negative = K.pack.n_negative;
    return PSS_OK;
}

// Whole-range streaming.  The range is cut into position chunks; chunk k + 1's reads (and reference slice)
// go up on their own stream into the second set of device buffers while chunk k is computed, copied
// back and widened, so the host -> device copies are hidden behind the (longer) result path.
extern "C" int pss_pileup_host_stream(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats, const pss_pileup_params *P,
                                      int32_t *const *outs, uint8_t *const *rowmasks, int64_t max_span,
                                      int64_t chunk_positions) {
    if (!ctx) return PSS_EINVAL;
    int rc;
    bool need_qual = false, need_seq = false;
    if ((rc = check_multi(ctx, reads, n_stats, P, outs, rowmasks, &need_qual, &need_seq)) != PSS_OK) return rc;
    const int64_t start = P[0].start, end = P[0].end, npos = end - start;
    if (chunk_positions <= 0) chunk_positions = 25000000;
    chunk_positions = (chunk_positions + PSS_TTILE - 1) / PSS_TTILE * PSS_TTILE;
    // name_prev indexes the whole batch, which a sub-range view cannot express: one call then
    if (need_qual || max_span < 1 || npos <= chunk_positions || reads->n_reads == 0)
        return pss_pileup_host_multi(ctx, reads, n_stats, P, outs, rowmasks);
    {
        // The depth cap (bam_plp_push) is order dependent: the first reads of a chunk would not see
        // the live reads of the previous one.  A read can only be refused when max_depth earlier reads start within
        // one maximal read span before it (the test of k_safe, on the host); if any read could be, the range is
        // not cut -- one call over the whole batch, whose admission pass sees every read in file order.
        const int64_t nr = reads->n_reads, maxcnt = P[0].max_depth ? P[0].max_depth : 8000;
        const int32_t *ps = reads->pos;
        bool refusable = false;
        for (int64_t i = maxcnt; i < nr && !refusable; ++i) refusable = (int64_t)ps[i] - (int64_t)ps[i - maxcnt] <= max_span;
        if (refusable) return pss_pileup_host_multi(ctx, reads, n_stats, P, outs, rowmasks);
    }
    for (int i = 0; i < n_stats; ++i)
        if (P[i].stat == PSS_COVERAGE_GC) return pss_pileup_host_multi(ctx, reads, n_stats, P, outs, rowmasks);
    CU(cudaSetDevice(ctx->device));
    ctx->prepared = false;
    if (!ctx->up_stream) {
        CU(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t s = nullptr;
    const int64_t n = reads->n_reads;
    const int32_t *pos = reads->pos;
    auto lower = [&](int64_t key) {     // first read index with pos >= key
        int64_t lo = 0, hi = n;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if ((int64_t)pos[mid] < key) lo = mid + 1; else hi = mid; }
        return lo;
    };
    const int64_t nchunks = (npos + chunk_positions - 1) / chunk_positions;
    const uint8_t *ref_host = nullptr;
    int64_t ref_start = 0, ref_len = 0;
    for (int i = 0; i < n_stats && !ref_host; ++i)
        if (P[i].ref && P[i].ref_len > 0) { ref_host = P[i].ref; ref_start = P[i].ref_start; ref_len = P[i].ref_len; }
    pss_reads Hk[2], Dk[2];
    const uint8_t *refd[2] = {nullptr, nullptr};
    int64_t refs[2] = {0, 0}, refl[2] = {0, 0};
    auto upload = [&](int64_t k) -> int {        // enqueue the upload of chunk k on up_stream
        const int b = (int)(k & 1);
        const int64_t c_lo = start + k * chunk_positions, c_hi = std::min(end, c_lo + chunk_positions);
        const int64_t i0 = lower(c_lo - max_span + 1), i1 = lower(c_hi);
        pss_reads v = *reads;
        v.n_reads = i1 - i0;
        v.pos += i0; v.flag += i0; v.mapq += i0; v.tlen += i0; v.mtid += i0; v.l_seq += i0;
        v.cigar_off += i0; v.seq_off += i0;
        Hk[b] = v;
        DevBuf *hb = b ? ctx->hb2 : ctx->hb;
        if (k >= 2) CU(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_free[b], 0));   // chunk k - 2 is done with these buffers
        int r = upload_reads(ctx, &v, &Dk[b], need_seq, false, ctx->up_stream, hb);
        if (r != PSS_OK) return r;
        refd[b] = nullptr;
        if (ref_host) {                          // the chunk's slice of the reference
            int64_t lo = std::max(ref_start, c_lo), hi = std::min(ref_start + ref_len, c_hi);
            if (hi > lo) {
                if ((r = to_device(ctx, hb[13], ref_host + (lo - ref_start), (size_t)(hi - lo), &refd[b], ctx->up_stream)) != PSS_OK) return r;
                refs[b] = lo; refl[b] = hi - lo;
            }
        }
        CU(cudaEventRecord(ctx->ev_up[b], ctx->up_stream));
        return PSS_OK;
    };
    if ((rc = upload(0)) != PSS_OK) return rc;
    std::vector<pss_pileup_params> Pc((size_t)n_stats);
    std::vector<int32_t *> oc((size_t)n_stats);
    std::vector<uint8_t *> mc((size_t)n_stats);
    for (int64_t k = 0; k < nchunks; ++k) {
        const int b = (int)(k & 1);
        const int64_t c_lo = start + k * chunk_positions, c_hi = std::min(end, c_lo + chunk_positions);
        if (k + 1 < nchunks && (rc = upload(k + 1)) != PSS_OK) return rc;
        CU(cudaStreamWaitEvent(s, ctx->ev_up[b], 0));
        for (int i = 0; i < n_stats; ++i) {
            Pc[(size_t)i] = P[i];
            Pc[(size_t)i].start = c_lo; Pc[(size_t)i].end = c_hi;
            if (P[i].ref && P[i].ref_len > 0) { Pc[(size_t)i].ref = refd[b] ? P[i].ref : nullptr; Pc[(size_t)i].ref_start = refs[b]; Pc[(size_t)i].ref_len = refd[b] ? refl[b] : 0; }
            oc[(size_t)i] = outs[i] + (c_lo - start);
            mc[(size_t)i] = rowmasks[i] + (c_lo - start);
        }
        if ((rc = host_multi_run(ctx, &Hk[b], &Dk[b], n_stats, Pc.data(), refd[b], false, oc.data(), npos, mc.data(), s)) != PSS_OK)
            return rc;
        CU(cudaEventRecord(ctx->ev_free[b], s));
        if ((rc = pss_synchronize(ctx, s)) != PSS_OK) return rc;       // unsorted / too long reads are reported per chunk
    }
    CU(cudaStreamSynchronize(ctx->up_stream));
    return PSS_OK;
}

extern "C" int pss_pileup_host(pss_ctx *ctx, const pss_reads *reads, const pss_pileup_params *P, int32_t *out,
                               uint8_t *rowmask) {
    if (!ctx) return PSS_EINVAL;
    if (!P) return fail(ctx, PSS_EINVAL, "params is NULL");
    return pss_pileup_host_multi(ctx, reads, 1, P, &out, &rowmask);
}

// ------------------------------------------------------------------------------ binned

static bool binned_needs_cigar(int stat) { return stat == PSS_COVERAGE_EXT_BINNED || stat == PSS_ALIGNMENT_BINNED; }
static bool binned_needs_ref(int stat) { return stat == PSS_COVERAGE_BINNED || stat == PSS_COVERAGE_EXT_BINNED; }

// the statistics of one call: each at most once, one common request; *mask receives the set (bit i = statistic i),
// by_stat[i] the caller's output pointer of statistic i
static int check_binned(pss_ctx *ctx, const pss_reads *R, int32_t n_stats, const pss_binned_params *P, int32_t *const *outs,
                        int *mask, int32_t **by_stat) {
    if (n_stats < 1 || !P || !outs) return fail(ctx, PSS_EINVAL, "n_stats < 1 or a NULL array");
    if (!R) return fail(ctx, PSS_EINVAL, "reads is NULL");
    if (R->n_reads < 0 || R->n_reads > 0x7fffffff) return fail(ctx, PSS_EINVAL, "n_reads out of range");
    if (R->n_reads > 0 && (!R->pos || !R->flag || !R->mapq || !R->tlen || !R->mtid))
        return fail(ctx, PSS_EINVAL, "a required read array is NULL");
    *mask = 0;
    for (int i = 0; i < PSS_N_BINNED_STATS; ++i) by_stat[i] = nullptr;
    for (int i = 0; i < n_stats; ++i) {
        const pss_binned_params &Q = P[i];
        if (Q.stat < 0 || Q.stat >= PSS_N_BINNED_STATS) return fail(ctx, PSS_EINVAL, "unsupported statistics type: %d", Q.stat);
        if (Q.window_size <= 0) return fail(ctx, PSS_EINVAL, "window_size <= 0");
        if (Q.n_bins != pss_binned_nbins(Q.start, Q.end, Q.window_size)) return fail(ctx, PSS_EINVAL, "n_bins mismatch");
        if (Q.n_bins > 0x7fffffff) return fail(ctx, PSS_EINVAL, "too many windows for one call");
        if (Q.start != P[0].start || Q.end != P[0].end || Q.window_size != P[0].window_size || Q.min_mapq != P[0].min_mapq ||
            (Q.no_dup != 0) != (P[0].no_dup != 0))
            return fail(ctx, PSS_EINVAL, "statistics of one call must share start, end, window_size, min_mapq, no_dup");
        if (*mask >> Q.stat & 1) return fail(ctx, PSS_EINVAL, "statistic %d requested twice", Q.stat);
        if (R->n_reads > 0 && binned_needs_cigar(Q.stat) && (!R->cigar || !R->cigar_off)) return fail(ctx, PSS_EINVAL, "cigar arrays are NULL");
        if (binned_needs_ref(Q.stat) && !Q.ref) return fail(ctx, PSS_EINVAL, "ref is NULL");
        if (!outs[i]) return fail(ctx, PSS_EINVAL, "out is NULL");
        *mask |= 1 << Q.stat;
        by_stat[Q.stat] = outs[i];
    }
    return PSS_OK;
}

// the common request of a call, with the reference span of its first statistic that has one
static pss_binned_params binned_common(int32_t n_stats, const pss_binned_params *P) {
    pss_binned_params C = P[0];
    C.ref = nullptr; C.ref_start = 0; C.ref_len = 0;
    for (int i = 0; i < n_stats; ++i)
        if (binned_needs_ref(P[i].stat) && P[i].ref) { C.ref = P[i].ref; C.ref_start = P[i].ref_start; C.ref_len = P[i].ref_len; break; }
    return C;
}

static int run_binned(pss_ctx *ctx, const pss_reads *R, const pss_binned_params &C, int mask, int32_t *const *by_stat,
                      cudaStream_t s) {
    int rc;
    if ((rc = ensure(ctx, ctx->bwork, binned_work_bytes(mask, C.n_bins))) != PSS_OK) return rc;
    BinnedEvents ev = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    timing_pair(ctx, PSS_KTAG_BINNED_ACC, &ev.acc0, &ev.acc1);
    timing_pair(ctx, PSS_KTAG_BINNED_FIN, &ev.fin0, &ev.fin1);
    if (mask & ((1 << PSS_COVERAGE_BINNED) | (1 << PSS_COVERAGE_EXT_BINNED))) timing_pair(ctx, PSS_KTAG_BINNED_GC, &ev.gc0, &ev.gc1);
    CU(launch_binned_multi(*R, C, mask, by_stat, ctx->bwork.p, s, &ctx->launches, ctx->timing ? &ev : nullptr));
    return PSS_OK;
}

extern "C" int pss_binned_multi(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats, const pss_binned_params *P,
                                int32_t *const *outs, void *stream) {
    if (!ctx) return PSS_EINVAL;
    int rc, mask;
    int32_t *by_stat[PSS_N_BINNED_STATS];
    if ((rc = check_binned(ctx, reads, n_stats, P, outs, &mask, by_stat)) != PSS_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    return run_binned(ctx, reads, binned_common(n_stats, P), mask, by_stat, (cudaStream_t)stream);
}

extern "C" int pss_binned(pss_ctx *ctx, const pss_reads *reads, const pss_binned_params *P, int32_t *out, void *stream) {
    if (!ctx) return PSS_EINVAL;
    if (!P) return fail(ctx, PSS_EINVAL, "params is NULL");
    return pss_binned_multi(ctx, reads, 1, P, &out, stream);
}

extern "C" int pss_binned_host_multi(pss_ctx *ctx, const pss_reads *H, int32_t n_stats, const pss_binned_params *P,
                                     int32_t *const *outs) {
    if (!ctx) return PSS_EINVAL;
    int rc, mask;
    int32_t *by_stat[PSS_N_BINNED_STATS];
    if ((rc = check_binned(ctx, H, n_stats, P, outs, &mask, by_stat)) != PSS_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = nullptr;
    pss_reads D = *H;
    const size_t n = (size_t)H->n_reads;
    bool need_cigar = false;
    for (int i = 0; i < n_stats; ++i) need_cigar = need_cigar || binned_needs_cigar(P[i].stat);
    // the record fields this statistic set reads, once for all of them
#define UP(idx, field, count) if ((rc = to_device(ctx, ctx->hb[idx], H->field, (count), &D.field, s)) != PSS_OK) return rc
    UP(0, pos, n); UP(1, flag, n); UP(2, mapq, n);
    if (mask & ((1 << PSS_COVERAGE_EXT_BINNED) | (1 << PSS_TLEN_BINNED))) UP(3, tlen, n);
    if (mask & (1 << PSS_COVERAGE_EXT_BINNED)) UP(4, mtid, n);
    if (need_cigar && n > 0) {
        UP(6, cigar_off, n + 1);
        if ((rc = to_device_span(ctx, ctx->hb[7], H->cigar, H->cigar_off, n, &D.cigar, s)) != PSS_OK) return rc;
    }
#undef UP
    pss_binned_params C = binned_common(n_stats, P);
    if (C.ref && C.ref_len > 0) {
        if ((rc = to_device(ctx, ctx->hb[13], C.ref, (size_t)C.ref_len, &C.ref, s)) != PSS_OK) return rc;
    }
    // all result columns in one device buffer, one copy back per statistic
    size_t total = 0, off[PSS_N_BINNED_STATS];
    int32_t *dev_out[PSS_N_BINNED_STATS];
    for (int t = 0; t < PSS_N_BINNED_STATS; ++t) {
        off[t] = total;
        if (mask >> t & 1) total += (size_t)kBinnedFields[t] * (size_t)C.n_bins;
    }
    if ((rc = ensure(ctx, ctx->outbuf[0], sizeof(int32_t) * total)) != PSS_OK) return rc;
    for (int t = 0; t < PSS_N_BINNED_STATS; ++t) dev_out[t] = (mask >> t & 1) ? (int32_t *)ctx->outbuf[0].p + off[t] : nullptr;
    if ((rc = run_binned(ctx, &D, C, mask, dev_out, s)) != PSS_OK) return rc;
    for (int t = 0; t < PSS_N_BINNED_STATS; ++t) {
        if (!(mask >> t & 1)) continue;
        const size_t obytes = sizeof(int32_t) * (size_t)kBinnedFields[t] * (size_t)C.n_bins;
        CU(cudaMemcpyAsync(by_stat[t], dev_out[t], obytes, cudaMemcpyDeviceToHost, s));
        ctx->d2h_bytes += (int64_t)obytes;
    }
    CU(cudaStreamSynchronize(s));
    return PSS_OK;
}

extern "C" int pss_binned_host(pss_ctx *ctx, const pss_reads *H, const pss_binned_params *P, int32_t *out) {
    if (!ctx) return PSS_EINVAL;
    if (!P) return fail(ctx, PSS_EINVAL, "params is NULL");
    return pss_binned_host_multi(ctx, H, 1, P, &out);
}
