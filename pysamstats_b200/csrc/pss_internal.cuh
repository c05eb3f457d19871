// Internal declarations shared by the CUDA translation units of libpss_b200.
// Not part of the public ABI (that is include/pss_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pss_b200.h"

// ---------------------------------------------------------------------------------
// Per-read record produced by k_prep and consumed by every later kernel.  32 bytes,
// so a tile's candidate reads are one contiguous, 16-byte aligned span in HBM.
// ---------------------------------------------------------------------------------
struct __align__(16) RInfo {
    int32_t pos;        // leftmost reference position
    int32_t end;        // htslib bam_endpos(): pos + reference length (>= pos + 1)
    uint32_t cls;       // CLS_* bits below
    uint32_t n_cigar;
    int64_t cigar_off;  // word index into the batch's cigar array
    int64_t seq_off;    // byte index into the batch's seq4 array
};

enum : uint32_t {
    CLS_REV = 1u << 0,        // BAM_FREVERSE
    CLS_PP = 1u << 1,         // BAM_FPROPER_PAIR
    CLS_ADMITTED = 1u << 2,   // enters htslib's pileup list (stepper passed, mapped, not capped)
    CLS_COUNTED = 1u << 3,    // ADMITTED and passes the read-level filters min_mapq / no_dup
    CLS_EXT_SHIFT = 4,        // bits 4..11: all, pp, mate_unmapped, mate_other_chr,
                              //             mate_same_strand, faceaway, softclipped, duplicate
    CLS_PAIRED = 1u << 12,    // tlen "paired" class: mate mapped on the same contig
    CLS_SIMPLE = 1u << 13,    // exactly one CIGAR op and it is M / = / X
    CLS_HASSEQ = 1u << 14,    // l_seq > 0
    CLS_NOREF = 1u << 15,     // CIGAR consumes no reference (end == pos + 1 by convention)
    CLS_MAPQ_SHIFT = 16,      // bits 16..23: mapping quality
    CLS_PUSHDEL = 1u << 24,   // reaches bam_plp_push but is ignored there (unmapped, nofilter)
    CLS_DROPPED = 1u << 25,   // refused by the max_depth rule
    CLS_OVL = 1u << 26,       // qualifies for the mate-overlap hash (A.3)
};

// BAM CIGAR op codes
enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };
__host__ __device__ __forceinline__ bool op_is_match(int op) { return op == OP_M || op == OP_EQ || op == OP_X; }

#define PSS_TILE 256          // positions per CTA in the pileup kernel (one per thread)
#define PSS_CHUNK 128         // candidate reads staged in shared memory at a time
#ifndef PSS_TTILE
#define PSS_TTILE 512         // positions per CTA in the tile (scatter) kernel
#endif

// device scalars shared between the small kernels
struct SmallState {
    int32_t maxspan;       // max over reads of end - pos
    int32_t unsorted;      // != 0 when pos[] decreases somewhere
    int32_t first_unsafe;  // depth-cap: first / last read index whose admission is not
    int32_t last_unsafe;   //            provably unconstrained (else INT_MAX / -1)
    int32_t toolong;       // != 0 when a read's sequence does not fit the tile kernel's staging buffer
    int32_t overflow16;    // != 0 when a tile needed several passes while writing 16-bit output
    int32_t pad_[2];
};

struct PileupArgs {
    const RInfo *ri;
    const uint32_t *cigar;
    const uint8_t *seq4;
    const uint8_t *qual;        // tweaked qualities (workspace copy) or NULL
    const int64_t *qual_off;
    const int64_t *seq_off;
    const int32_t *tlen;
    const int2 *tile_rng;       // candidate read range [x, y) per tile
    const uint32_t *lowq;       // tile kernel, min_baseq > 0: bit per base below the threshold (k_lowq)
    int32_t *redo;              // tlen: the in-order replay list -- [0] = number of 256-position tiles to redo, [1..] their
                                // indices: k_tilet appends the halves of a 512-position tile it cannot decide, the
                                // replay launch of k_pileup (a small fixed grid) walks the list
    int64_t n_reads;
    const uint8_t *ref;
    int64_t ref_start, ref_len;
    int64_t start, end;         // output position range
    int32_t min_baseq;
    int32_t no_del;
    const SmallState *st_ro;    // device scalars (maxspan)
    int32_t *toolong;           // set when a read does not fit the staging buffers
    int32_t *out;
    int64_t out_stride;         // elements between consecutive fields of out (>= end - start)
    int32_t narrow;             // tile kernel only: out holds uint16 (1) / uint8 (2) elements (host transport)
    int32_t *ovf8;              // narrow == 2: set when a count does not fit 8 bits
    uint8_t *rowmask;
};


#ifdef __CUDACC__
// ------------------------------------------------------------------ async-copy plumbing

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// ask for [src, src + bytes) (16-byte aligned, multiple of 16) to be brought into L2
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Bulk copies need 16-byte aligned addresses and sizes: widen the element span [lo, hi) of
// `base` outward to 16-byte boundaries.  *first = element index that lands at offset 0 of
// the shared buffer (may be fractional for 1-byte elements only in the sense of bytes).
template <class Tp>
__device__ __forceinline__ void stage_span(const Tp *base, long long lo, long long hi, int elem, const uint8_t **src,
                                           uint32_t *bytes, long long *first) {
    uintptr_t a0 = reinterpret_cast<uintptr_t>(base + lo), a1 = reinterpret_cast<uintptr_t>(base + hi);
    uintptr_t al = a0 & ~(uintptr_t)15, ah = (a1 + 15) & ~(uintptr_t)15;
    *src = reinterpret_cast<const uint8_t *>(al);
    *bytes = (uint32_t)(ah - al);
    *first = lo - (long long)((a0 - al) / (uintptr_t)elem);
}

#endif  // __CUDACC__

// ---- kernel launchers (defined in the .cu files) ----------------------------------
cudaError_t launch_prep(const pss_reads &R, int stepper, int min_mapq, int no_dup, RInfo *ri,
                        SmallState *st, cudaStream_t s, int64_t *launches);
cudaError_t launch_admit(RInfo *ri, const int32_t *pos, int64_t n, int maxcnt, SmallState *st, cudaStream_t s,
                         int64_t *launches);
cudaError_t launch_overlap(const pss_reads &R, const RInfo *ri, uint8_t *qual_work, int32_t *pair_first,
                           int32_t *name_next, cudaStream_t s, int64_t *launches);
cudaError_t launch_tile_ranges(const RInfo *ri, int64_t n, const SmallState *st, int64_t start,
                               int64_t ntiles, int tile, int2 *rng, cudaStream_t s, int64_t *launches);
bool tile_kernel_supports(int stat, int min_baseq, int no_del);
bool tile_kernel_counting(int stat);
size_t tile_desc_bytes(int64_t ntiles);
size_t tile_redo_bytes(int64_t ntiles);
size_t tile_lowq_bytes(int64_t n_reads, int64_t n_seq_bytes);
cudaError_t launch_lowq(const pss_reads &R, const uint8_t *qual, int min_baseq, void *lowq_buf, cudaStream_t s, int64_t *launches);
size_t tile_refpack_bytes(int64_t ntiles);
// e0 / e1 (may be NULL) are recorded immediately before / after the tile kernel itself
cudaError_t launch_tile(int stat, const PileupArgs &p, const int64_t *seq_off, int64_t n_reads, SmallState *st,
                        int need_raw, int64_t ntiles, void *desc_buf, void *refpack_buf, bool desc_valid,
                        cudaStream_t s, int64_t *launches, cudaEvent_t e0, cudaEvent_t e1);
cudaError_t launch_pileup(int stat, const PileupArgs &a, int64_t ntiles, bool staged, cudaStream_t s, int64_t *launches);
cudaError_t launch_gc(const uint8_t *ref, int64_t ref_start, int64_t ref_len, int64_t contig_len,
                      int64_t start, int64_t end, int window_size, int window_offset, int32_t *out_gc,
                      cudaStream_t s, int64_t *launches);
// binned statistics (pss_kernels_binned.cu): `mask` bit i = enum pss_binned_stat i
struct BinnedEvents { cudaEvent_t acc0, acc1, fin0, fin1, gc0, gc1; };   // recorded around the kernels when not NULL
size_t binned_work_bytes(int mask, int64_t n_bins);
cudaError_t launch_binned_multi(const pss_reads &R, const pss_binned_params &P, int mask, int32_t *const *outs, void *work,
                                cudaStream_t s, int64_t *launches, const BinnedEvents *ev);
cudaError_t launch_export_admitted(const RInfo *ri, int64_t n, uint8_t *admitted, cudaStream_t s,
                                   int64_t *launches);
