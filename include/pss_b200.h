/*
 * pss_b200.h -- C ABI of the B200-native pileup / binned statistics engine.
 *
 * This is the drop-in boundary for the hot path of alimanfoo/pysamstats: everything the
 * reference computes inside pysamstats/opt.pyx (+ the htslib pileup engine it drives) for
 * one contig region, expressed over plain pointers and sizes.  No torch / numpy types.
 * The python host side (pysamstats_b200/_cabi.py) binds these with ctypes; see
 * INTEGRATION.md for the binding a pysamstats maintainer would add.
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative PSS_E* code; the message is
 *     available through pss_last_error().  Nothing throws across the ABI.
 *   - the caller owns every buffer passed in.  "dev" pointers are CUDA device pointers
 *     (e.g. torch tensors' data_ptr()), "host" pointers are ordinary (preferably pinned)
 *     host memory.  The context owns only its private scratch workspace.
 *   - one pss_ctx per (GPU, host thread); no global state.  Work is launched on the
 *     stream given (a cudaStream_t passed as void*, NULL = default stream); the *_host
 *     entry points synchronise before returning, the device entry points do not.
 *   - a read batch is the set of BAM records of ONE contig in file (coordinate) order,
 *     as structure-of-arrays.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef PSS_B200_H
#define PSS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSS_OK 0
#define PSS_EINVAL (-1)   /* bad argument */
#define PSS_ECUDA (-2)    /* CUDA runtime / launch failure */
#define PSS_ENOMEM (-3)   /* workspace allocation failed */
#define PSS_EUNSORTED (-4) /* batch not coordinate sorted (htslib: "unsorted positions") */
#define PSS_ETOOLONG (-5)  /* a read of several kb met the short-read kernels; see PSS_FLAG_LONG_READS (the *_host
                              entry points repeat such a call with that flag themselves) */

/* pss_pileup_params.flags */
#define PSS_FLAG_GENERIC_KERNEL 1 /* always use the generic per-position kernel (testing / comparison) */
#define PSS_FLAG_LONG_READS 2     /* reads of arbitrary length: per-position kernel without shared-memory staging */

/* statistics ids: order of stats_classes_pileup, pysamstats/pileup.py:150-166 */
enum pss_pileup_stat {
    PSS_COVERAGE = 0,
    PSS_COVERAGE_STRAND = 1,
    PSS_COVERAGE_EXT = 2,
    PSS_COVERAGE_EXT_STRAND = 3,
    PSS_VARIATION = 4,
    PSS_VARIATION_STRAND = 5,
    PSS_TLEN = 6,
    PSS_TLEN_STRAND = 7,
    PSS_MAPQ = 8,
    PSS_MAPQ_STRAND = 9,
    PSS_BASEQ = 10,
    PSS_BASEQ_STRAND = 11,
    PSS_BASEQ_EXT = 12,
    PSS_BASEQ_EXT_STRAND = 13,
    PSS_COVERAGE_GC = 14,
    PSS_N_PILEUP_STATS = 15
};

/* order of stats_classes_binned, pysamstats/binned.py:121-127 */
enum pss_binned_stat {
    PSS_COVERAGE_BINNED = 0,
    PSS_COVERAGE_EXT_BINNED = 1,
    PSS_MAPQ_BINNED = 2,
    PSS_ALIGNMENT_BINNED = 3,
    PSS_TLEN_BINNED = 4,
    PSS_N_BINNED_STATS = 5
};

/* pysam stepper names, pysamstats/config.py:29-31 */
enum pss_stepper { PSS_STEPPER_ALL = 0, PSS_STEPPER_NOFILTER = 1, PSS_STEPPER_SAMTOOLS = 2 };

typedef struct pss_ctx pss_ctx;

/*
 * Structure-of-arrays read batch: the fields of htslib's bam1_t that opt.pyx reads
 * (opt.pyx:10-14: core.pos/flag/qual/isize/mtid/n_cigar/l_qseq, bam_get_cigar/seq/qual).
 * Offsets arrays have n_reads + 1 entries and index the cigar / seq4 / qual pointers given
 * here; they need not start at 0, so a sub-range of a larger batch is a zero-copy view
 * (advance the per-read pointers, keep the payload pointers).  The *_host entry points read
 * only [off[0], off[n_reads]) of each payload and ignore n_cigar_words / n_seq_bytes /
 * n_qual_bytes; the device entry points require qual_off[0] == 0 and n_qual_bytes set.
 * seq4 is BAM's packed 4-bit sequence ((l_seq + 1) / 2 bytes per read), qual is l_seq
 * bytes per read.
 * qual / qual_off / name_prev may be NULL when the request does not depend on base
 * qualities (no baseq* statistic and min_baseq == 0).
 * name_prev[i] = index of the previous record in this batch with the same QNAME, or -1:
 * the QNAME equivalence htslib's mate-overlap hash keys on.
 */
typedef struct pss_reads {
    int64_t n_reads;
    int64_t n_cigar_words; /* elements in cigar[]  (== cigar_off[n_reads]) */
    int64_t n_seq_bytes;   /* bytes in seq4[]      (== seq_off[n_reads]) */
    int64_t n_qual_bytes;  /* bytes in qual[]      (== qual_off[n_reads]), 0 if qual is NULL */
    int32_t tid;           /* contig id of the batch (col.tid in opt.pyx:272) */
    int32_t reserved;
    const int32_t *pos;   /* core.pos, non-decreasing */
    const uint16_t *flag; /* core.flag */
    const uint8_t *mapq;  /* core.qual */
    const int32_t *tlen;  /* core.isize */
    const int32_t *mtid;  /* core.mtid */
    const int32_t *l_seq; /* core.l_qseq */
    const int64_t *cigar_off;
    const uint32_t *cigar; /* BAM encoding: len << 4 | op */
    const int64_t *seq_off;
    const uint8_t *seq4;
    const int64_t *qual_off;
    const uint8_t *qual;
    const int32_t *name_prev;
} pss_reads;

/*
 * Pileup request: the keyword arguments of opt.iter_pileup (opt.pyx:1827, :1846-1847) that
 * reach the engine.  [start, end) is the OUTPUT position range: out column j describes
 * position start + j.  (The host decides that range from region / truncate / pad and the
 * fetched reads; the engine does not clip reads to it.)
 * ref / ref_start / ref_len describe reference bytes (any case) for positions
 * [ref_start, ref_start + ref_len); positions outside behave like an empty
 * FastaFile.fetch() (opt.pyx:1987-1994).  ref may be NULL for statistics without `ref`.
 */
typedef struct pss_pileup_params {
    int32_t stat;          /* enum pss_pileup_stat */
    int32_t stepper;       /* enum pss_stepper */
    int32_t max_depth;     /* 0 -> htslib default 8000 */
    int32_t min_mapq;
    int32_t min_baseq;
    int32_t no_del;
    int32_t no_dup;
    int32_t window_size;   /* coverage_gc only (opt.pyx:1477-1481) */
    int32_t window_offset; /* coverage_gc only, already defaulted */
    int32_t flags;         /* PSS_FLAG_* */
    int64_t start;
    int64_t end;
    const uint8_t *ref;
    int64_t ref_start;
    int64_t ref_len;
    int64_t contig_len;    /* FASTA length of the contig (clips the coverage_gc window) */
} pss_pileup_params;

/* Binned request: keyword arguments of opt.iter_binned_chrom (opt.pyx:2026-2028). */
typedef struct pss_binned_params {
    int32_t stat;          /* enum pss_binned_stat */
    int32_t window_size;
    int32_t min_mapq;
    int32_t no_dup;
    int64_t start;         /* normalised region start (bin 0 starts here) */
    int64_t end;
    int64_t n_bins;        /* from pss_binned_nbins() */
    const uint8_t *ref;    /* coverage / coverage_ext: reference bytes, else NULL */
    int64_t ref_start;
    int64_t ref_len;
} pss_binned_params;

/* ---- life cycle ------------------------------------------------------------------ */

/* Creates a context on CUDA device `device`.  Replaces nothing in the reference (which
 * has no device); one context per GPU/thread. */
int pss_create(int device, pss_ctx **out);
void pss_destroy(pss_ctx *ctx);
const char *pss_last_error(const pss_ctx *ctx);
const char *pss_version(void);
/* Waits for `stream` and reports errors detected on the device by earlier asynchronous
 * calls on this context (PSS_EUNSORTED: the batch was not coordinate sorted; htslib's
 * pileup refuses such input too).  The device entry points never synchronise themselves. */
int pss_synchronize(pss_ctx *ctx, void *stream);
/* Number of CUDA kernels this context has launched so far (bench.py's gpu_launches). */
int64_t pss_launch_count(const pss_ctx *ctx);
/* Bytes the *_host entry points have copied host->device / device->host so far
 * (bench.py's e2e.h2d_bytes_per_step / d2h_bytes_per_step). */
void pss_transfer_bytes(const pss_ctx *ctx, int64_t *h2d, int64_t *d2h);
/* Average device time (ms, CUDA events on the launch stream) of the pileup kernel of
 * statistic `stat` (enum pss_pileup_stat; or one of the PSS_KTAG_* kernels below; -1 = any) over
 * its launches since the last call; those launches are then forgotten.  Used by bench.py for
 * roofline.achieved.  Returns < 0 if there were none. */
#define PSS_KTAG_GC 100          /* k_gc: the sliding-window GC% column of coverage_gc */
#define PSS_KTAG_PREP 101        /* k_prep + the depth-cap kernels (one preparation of a batch) */
#define PSS_KTAG_BINNED_ACC 110  /* k_binned_acc: the one pass over the read records of a binned call */
#define PSS_KTAG_BINNED_FIN 111  /* k_binned_fin */
#define PSS_KTAG_BINNED_GC 112   /* k_binned_gc */
double pss_pileup_kernel_ms(pss_ctx *ctx, int32_t stat, int64_t *n_launches);
/* Enable/disable the per-launch event timing above (default off: no overhead). */
void pss_set_kernel_timing(pss_ctx *ctx, int enabled);

/* ---- shapes ---------------------------------------------------------------------- */

/* number of int32 value fields per row (config.py dtype lists minus chrom/pos/ref) */
int32_t pss_pileup_nfields(int32_t stat);
int32_t pss_binned_nfields(int32_t stat);
/* 1 when the statistic carries a `ref` column / needs reference bytes */
int32_t pss_pileup_needs_ref(int32_t stat);
/* 1 when the request reads base qualities (so qual/qual_off/name_prev are required) */
int32_t pss_pileup_needs_qual(int32_t stat, int32_t min_baseq);
/* positions per tile of the scatter kernel: callers that cut a range themselves (shards, chunks) align interior
 * cuts to a multiple of this relative to `start`, so that no piece begins with a partial tile */
int32_t pss_tile_positions(void);
/* number of windows iter_binned_chrom yields for [start, end) (opt.pyx:2043-2102) */
int64_t pss_binned_nbins(int64_t start, int64_t end, int32_t window_size);

/* ---- pileup statistics ------------------------------------------------------------ */

/*
 * Replaces opt.iter_pileup_default / stat_pileup (opt.pyx:1846-1919) and the htslib
 * column builder behind AlignmentFile.pileup (call site opt.pyx:1853-1854) for one
 * contig: for every position p in [start, end)
 *     out[f * (end - start) + (p - start)] = field f of the record for p
 *     rowmask[p - start] = 1 iff htslib would emit a column at p (>= 1 read admitted,
 *                          before pysamstats' own filters -- opt.pyx:1892-1915)
 * Field order = pysamstats/config.py dtype_<stat> minus chrom/pos/ref.  Rows with
 * rowmask 0 hold the values of an empty accumulator (what pad=True emits).
 * All pointers in `reads`, params->ref, out and rowmask are DEVICE pointers.
 */
int pss_pileup(pss_ctx *ctx, const pss_reads *reads, const pss_pileup_params *params,
               int32_t *out, uint8_t *rowmask, void *stream);

/*
 * pss_pileup() in two halves, for callers that want several statistics of one batch.
 * pss_prepare() does everything that depends on the batch and on the read-level arguments
 * only (stepper flag filter, max_depth admission rule -- call site opt.pyx:1853-1854 --, the
 * read filters min_mapq / no_dup of opt.pyx:1897-1909 and, if with_qual, the mate-overlap
 * quality rewrite); pss_pileup_prepared() then computes one statistic over one position
 * range from the prepared batch.  Its params must carry the same stepper / max_depth /
 * min_mapq / no_dup.  The arrays of `reads` must stay valid and unchanged until the last
 * pss_pileup_prepared() call has completed; any other entry point of this context
 * discards the prepared batch.  DEVICE pointers.
 */
int pss_prepare(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth, int32_t min_mapq,
                int32_t no_dup, int32_t with_qual, void *stream);
int pss_pileup_prepared(pss_ctx *ctx, const pss_pileup_params *params, int32_t *out, uint8_t *rowmask,
                        void *stream);

/*
 * Same contract with HOST pointers everywhere: copies the batch to the device, runs
 * pss_pileup over position chunks and copies results back (synchronous).  This is the
 * call timed as `e2e` in bench.py.
 */
int pss_pileup_host(pss_ctx *ctx, const pss_reads *reads, const pss_pileup_params *params,
                    int32_t *out, uint8_t *rowmask);

/*
 * Several statistics of one batch and one position range in one call: the batch is copied
 * to the device and prepared once.  params[i] / outs[i] / rowmasks[i] describe statistic i;
 * all params must share stepper, max_depth, min_mapq, no_dup, start, end and the reference
 * span.  HOST pointers (synchronous).
 */
int pss_pileup_host_multi(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats,
                          const pss_pileup_params *params, int32_t *const *outs,
                          uint8_t *const *rowmasks);

/*
 * The same statistic delivered as packed RECORDS -- what load_* returns (util.load_stats,
 * /root/reference/pysamstats/util.py:31-73: one structured-array record per emitted genome
 * position, field order and packing of config.py:33-402).  The result chunks are packed from the
 * pinned staging buffers straight into the caller's records by the worker threads, compacted to the
 * rows the reference emits: a position gets a row iff htslib builds a column there (rowmask) or it
 * lies in [pad_start, pad_end) (pad=True, opt.pyx:1936-1984).  field_off[j] < 0 leaves value field j
 * out (fields=...); field_bytes[j] is 4 (int32) or 1 (low byte: coverage_gc's gc is u1).  chrom
 * (chrom_bytes bytes, zero padded by the caller), pos (int32, position + pos_add) and ref (one byte,
 * the upper-cased base of params->ref, 0 outside it: opt.pyx:1987-1994) are written when their offset
 * is >= 0.  Rows are written from `rows` on, at most capacity_rows of them (PSS_EINVAL beyond);
 * *n_rows receives their number, *n_negative how many emitted values of field `check_field` were
 * negative (gc = -1 of an empty window, which the reference cannot store).  Records beyond *n_rows (inside the
 * capacity) are unspecified: an attempt that had to be repeated with wider columns may have written them.
 * HOST pointers, synchronous.
 */
typedef struct pss_row_layout {
    int32_t row_bytes;
    int32_t n_fields;              /* must equal pss_pileup_nfields(params->stat) */
    const int32_t *field_off;
    const int32_t *field_bytes;
    const uint8_t *chrom;
    int32_t chrom_off, chrom_bytes;
    int32_t pos_off, ref_off;
    int32_t check_field, reserved;
    int64_t pos_add;
    int64_t pad_start, pad_end;
} pss_row_layout;

int pss_pileup_host_rows(pss_ctx *ctx, const pss_reads *reads, const pss_pileup_params *params,
                         const pss_row_layout *layout, void *rows, int64_t capacity_rows,
                         int64_t *n_rows, int64_t *n_negative);

/*
 * The same over a long position range, streamed: the library cuts [start, end) into chunks of
 * `chunk_positions` positions (0: 25 M), and uploads the reads of chunk k + 1 -- those with
 * pos in (chunk start - max_span, chunk end), a sub-range of the start-sorted batch -- while
 * chunk k is computed, copied back and widened.  `max_span` >= max(bam_endpos - pos) over the
 * batch (the decoder knows it; a larger value only costs bandwidth).  outs[i] is
 * [n_fields][end - start] as for pss_pileup_host_multi.  The results are those of ONE
 * pss_pileup_host_multi call over the whole batch, always: requests that read base qualities
 * (the mate links of name_prev index the whole batch), coverage_gc, and batches in which the
 * order-dependent max_depth rule could refuse a read at all (max_depth reads start within
 * max_span of some read: a chunk would not see the live reads of the previous one) are not cut
 * but fall back to that one call.  HOST pointers, pinned for the copies to overlap (synchronous).
 */
int pss_pileup_host_stream(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats,
                           const pss_pileup_params *params, int32_t *const *outs,
                           uint8_t *const *rowmasks, int64_t max_span, int64_t chunk_positions);

/*
 * The depth-cap admission rule of bam_plp_push (max_depth, call site opt.pyx:1853-1854)
 * plus the stepper flag filter, exposed on its own: admitted[i] = 1 iff read i enters the
 * pileup.  DEVICE pointers.
 */
int pss_admit(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth,
              uint8_t *admitted, void *stream);

/*
 * The mate-overlap base-quality rewrite htslib applies under pysam's default
 * ignore_overlaps=True (call site opt.pyx:1853), exposed on its own:
 * qual_out receives the qualities the pileup sees (same layout as reads->qual),
 * pair_first[i] = index of the earlier mate read i was paired with, or -1.
 * DEVICE pointers; pair_first may be NULL.
 */
int pss_overlap_tweak(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth,
                      uint8_t *qual_out, int32_t *pair_first, void *stream);

/*
 * The PAIRS of that rewrite for a whole batch, from the read metadata alone: pair_first[i] as above, for
 * reads->n_reads reads, computed from pos / flag / cigar / name_prev (sequences and qualities are neither
 * needed nor copied; seq4 / qual may point anywhere).  This is what lets a caller cut a request that reads
 * base qualities into position pieces (the reference streams such a request through htslib's overlap hash,
 * call site opt.pyx:1853; the hash keys on QNAMEs, so which reads pair depends on reads a piece does not
 * hold): a piece passes name_prev[b] = a - first for every pair (a, b) whose reads both lie in the piece and
 * -1 everywhere else, and pairs exactly as the whole batch does.  HOST pointers, synchronous.
 */
int pss_pairs_host(pss_ctx *ctx, const pss_reads *reads, int32_t stepper, int32_t max_depth,
                   int32_t *pair_first);

/* ---- binned statistics ------------------------------------------------------------ */

/*
 * Replaces opt.iter_binned_chrom (opt.pyx:2026-2102) and the five BinnedStat classes
 * (opt.pyx:1551-1819) for one contig region: out[f * n_bins + k] = field f of window k,
 * window k = [start + k*W, start + (k+1)*W).  Field order = config.py dtype_<stat>_binned
 * minus chrom/pos.  `reads` must be the records IteratorRowRegion yields for the region
 * (pos < end and endpos > start, file order).  DEVICE pointers.
 */
int pss_binned(pss_ctx *ctx, const pss_reads *reads, const pss_binned_params *params,
               int32_t *out, void *stream);

/* Same with HOST pointers (synchronous). */
int pss_binned_host(pss_ctx *ctx, const pss_reads *reads, const pss_binned_params *params,
                    int32_t *out);

/*
 * Several binned statistics of one region in ONE pass over the read records -- the reference
 * runs iter_binned_chrom once per statistic (opt.pyx:2026-2102, one BinnedStat per call); the
 * read stream is what bounds the kernel, so a sweep such as coverage_binned + mapq_binned +
 * tlen_binned reads pos / flag / mapq / tlen once instead of three times.  params[i] / outs[i]
 * describe statistic i (layout of pss_binned); every statistic at most once, and all params
 * share start, end, window_size, n_bins, min_mapq, no_dup (the reference span is taken from the
 * first statistic that needs one).  DEVICE pointers / HOST pointers (synchronous).
 */
int pss_binned_multi(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats, const pss_binned_params *params,
                     int32_t *const *outs, void *stream);
int pss_binned_host_multi(pss_ctx *ctx, const pss_reads *reads, int32_t n_stats, const pss_binned_params *params,
                          int32_t *const *outs);

#ifdef __cplusplus
}
#endif
#endif /* PSS_B200_H */
