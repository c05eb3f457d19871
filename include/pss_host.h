/*
 * pss_host.h -- C ABI of the host-side BAM decoder / writer (libpss_host.so, plain C++ + zlib,
 * no CUDA).  It produces the structure-of-arrays read batch that include/pss_b200.h consumes.
 *
 * Replaces, for this engine, what the reference obtains from pysam / htslib:
 *   - BGZF inflate + BAM record parsing behind AlignmentFile.pileup (pysamstats/opt.pyx:1853)
 *     and IteratorRowRegion / cnext (opt.pyx:2040-2048);
 *   - bam_endpos() (used for the region rule pos < end && endpos > start);
 *   - the QNAME equivalence htslib's overlap hash keys on (pss_name_prev).
 * The python binding is pysamstats_b200/bamio.py.  All functions return 0 on success or a
 * negative code; pss_host_error() gives the message of the last failure on this thread.
 */
#ifndef PSS_HOST_H
#define PSS_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pss_bam pss_bam;

/* record arrays of a whole BAM file, file order; every pointer is caller-allocated with the
 * sizes reported by pss_bam_sizes() (offset arrays: n + 1 entries) */
typedef struct pss_host_arrays {
    int32_t *tid;
    int32_t *pos;
    uint16_t *flag;
    uint8_t *mapq;
    int32_t *tlen;
    int32_t *mtid;
    int32_t *mpos;
    int32_t *l_seq;
    int64_t *endpos;      /* htslib bam_endpos(): pos + reference length (pos + 1 if unmapped / zero) */
    int64_t *cigar_off;
    uint32_t *cigar;      /* BAM encoding: len << 4 | op */
    int64_t *seq_off;
    uint8_t *seq4;        /* packed 4-bit bases, (l_seq + 1) / 2 bytes per record */
    int64_t *qual_off;
    uint8_t *qual;
    uint8_t *names;       /* n x name_width bytes, NUL padded */
} pss_host_arrays;

const char *pss_host_error(void);

/* Reads the file, inflates all BGZF blocks on n_threads threads (<= 0: hardware concurrency)
 * and indexes the BAM records. */
int pss_bam_open(const char *path, int n_threads, pss_bam **out);
/* Region open through the BAI index (replaces pysam's IteratorRowRegion / hts_itr_query behind
 * alignmentfile.fetch / pileup with a region, opt.pyx:1853-1854, :2038-2041): reads and inflates
 * only the BGZF blocks the index lists for reference `tid`:[beg, end) and keeps the records with
 * pos < end && bam_endpos > beg, in file order.  bai_path NULL: "<path>.bai".  tid < 0: header
 * only (no index needed, no records). */
int pss_bam_open_region(const char *path, const char *bai_path, int32_t tid, int64_t beg, int64_t end,
                        int n_threads, pss_bam **out);
/* Writes the BAI index of a coordinate-sorted BAM (bai_path NULL: "<path>.bai"). */
int pss_bam_index_build(const char *path, const char *bai_path, int n_threads);
void pss_bam_close(pss_bam *b);

int32_t pss_bam_n_ref(const pss_bam *b);
const char *pss_bam_ref_name(const pss_bam *b, int32_t i);
int64_t pss_bam_ref_len(const pss_bam *b, int32_t i);

/* n records, total cigar words / packed sequence bytes / quality bytes, QNAME column width */
int pss_bam_sizes(const pss_bam *b, int64_t *n, int64_t *n_cigar, int64_t *n_seq, int64_t *n_qual,
                  int32_t *name_width);
/* fills the arrays (parallel over records) */
int pss_bam_fill(const pss_bam *b, const pss_host_arrays *a);
/* seconds spent in (file read, inflate, record index) by pss_bam_open and bytes inflated */
int pss_bam_timing(const pss_bam *b, double *t_read, double *t_inflate, double *t_index, int64_t *raw_bytes);

/* Packs columns into `n_rows` records of `row_bytes` bytes (the packed structured array load_* returns;
 * replaces the np.fromiter row assembly of pysamstats/util.py:61-71): column c contributes n_bytes[c] bytes
 * at offset dst_off[c] of every record, read from src[c] + row * src_stride[c] (stride 0: one constant value
 * for every row; stride 8 with n_bytes 4: the low half of a little-endian int64). */
int pss_pack_rows(uint8_t *dst, int64_t n_rows, int32_t row_bytes, int32_t n_cols, const void *const *src,
                  const int32_t *src_stride, const int32_t *n_bytes, const int32_t *dst_off, int n_threads);

/* dst[c * dst_stride + k] = src[c * src_stride + idx[k]] for n_cols int32 columns: compaction of the engine's
 * per-position columns to the emitted rows (strides in elements). */
int pss_take_cols_i32(const int32_t *src, int64_t src_stride, int32_t n_cols, const int64_t *idx, int64_t n,
                      int32_t *dst, int64_t dst_stride, int n_threads);

/* One raw DEFLATE stream (the payload of a BGZF block; htslib bgzf.c inflates these behind every pysam read,
 * opt.pyx:1853-1854, :2038-2048) inflated whole into dst[0, out_len): mode 0 = the decoder of the BAM path
 * (pss_inflate.cpp, zlib for whatever it declines), 1 = that decoder alone (-1 when it declines), 2 = zlib alone.
 * 0 on success.  For tests and timing. */
int pss_inflate_raw(const uint8_t *src, int64_t in_len, uint8_t *dst, int64_t out_len, int mode);

/* Text of n_rows rows as csv.writer would write them for values that need no quoting (replaces the per-row
 * csv.writer.writerow of pysamstats/io.py:60-75): column c is kind 0 = constant text (src = its bytes, width =
 * their number), 1 = int32, 2 = int64 (decimal), 3 = one-byte text (NUL = empty); src_stride in bytes; cells
 * joined by `delimiter`, rows ended by `terminator`.  *out is malloc'ed; release it with pss_host_free. */
int pss_format_rows(int64_t n_rows, int32_t n_cols, const int32_t *kind, const void *const *src,
                    const int32_t *src_stride, const int32_t *width, const char *delimiter, int32_t delimiter_len,
                    const char *terminator, int32_t terminator_len, int n_threads, char **out, int64_t *out_len);
void pss_host_free(void *p);
/* Host memory bandwidth probe (bench.py's e2e.host_floor): n_threads threads (0: all) stream `reps` times over a
 * private buffer of `bytes` each.  mode 0: non-temporal int32 stores (what the result path does to the caller's
 * int32 columns -- the API contract of load_* / pss_pileup_host*, /root/reference/pysamstats/util.py:61-71, fixes
 * how many of those bytes the host must write), 1: widen u8 -> int32, 2: reads.  Returns GB/s, < 0 on failure. */
double pss_host_bandwidth(int64_t bytes, int n_threads, int reps, int mode);

/* prev[i] = index of the previous row with the same (NUL padded, fixed width) name, or -1 */
int pss_name_prev(const uint8_t *names, int64_t n, int32_t width, int32_t *prev);

/* Writes a coordinate-sorted BAM (BGZF, deflate level `level`, blocks compressed on n_threads
 * threads) from SoA arrays of ONE batch: all records get `tid` from the tid array.  `names`
 * as above.  Used to materialise the synthetic workload as a real BAM file. */
int pss_bam_write(const char *path, int32_t n_ref, const char *const *ref_names, const int64_t *ref_lens,
                  int64_t n, const pss_host_arrays *a, int32_t name_width, int level, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* PSS_HOST_H */
