// The tile kernel: the per-read scatter path for the counting statistics (coverage,
// coverage_strand, coverage_gc, coverage_ext, coverage_ext_strand, variation,
// variation_strand -- pysamstats/opt.pyx:107-670, :1471-1519) when min_baseq == 0.
//
// Design (sm_100a):
//   * one CTA owns a tile of TT genome positions; its per-position counters live in shared
//     memory (counter-major, 4 bytes each) for the whole life of the tile;
//   * the tile's candidate reads are a contiguous index range of the start-sorted batch;
//     their 32-byte records and their packed 4-bit sequences are staged into shared memory
//     by bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier), chunk by
//     chunk for arbitrarily deep tiles;
//   * one thread per read walks the CIGAR once.  Everything that is constant along an
//     aligned span (read counts per strand / proper-pair / mate class, deletions, ref
//     skips) is an interval add: +1 at the span start, -1 at its end in a difference
//     array.  Bases are NOT visited one by one: the read's 4-bit codes are XOR-ed, eight
//     bases per 32-bit word, against the reference tile packed the same way, and only the
//     mismatching bases (<1 % of them) become point updates.  matches and the A/C/G/T/N
//     counters follow exactly from  matches = aligned - mismatches  and the reference code
//     of the position;
//   * after the last chunk the difference arrays are prefix-summed in place (warp shuffle
//     scan), each thread finalises its positions (all 60 fields of variation_strand are
//     linear in the counters) and writes the SoA columns with 128-byte coalesced stores.
// Shared-memory atomics are few (a handful per read), so the kernel is bound by the output
// stream to HBM.
#include "pss_internal.cuh"

#define FULL 0xffffffffu

#define TT PSS_TTILE          // positions per tile
#define TTHREADS 256
#define TCH 256               // reads staged per chunk (one thread each)
#define TSEQ_BYTES 20480      // staged sequence bytes per chunk (256 x 75 B fits)

enum { KIND_COV = 0, KIND_EXT = 1, KIND_VAR = 2 };

template <int KIND> struct KindInfo;
template <> struct KindInfo<KIND_COV> { static const int NDIFF = 4, NPOINT = 0; };    // all[bank]
template <> struct KindInfo<KIND_EXT> { static const int NDIFF = 16, NPOINT = 0; };   // class k * 2 + strand
// nondel[b], del[4 + b], skip[8 + b] | ins[b], mismatch class c in A,C,G,T,N,other: [4 + 4 c + b]
template <> struct KindInfo<KIND_VAR> { static const int NDIFF = 12, NPOINT = 28; };

struct TileArgs {
    const RInfo *ri;
    const uint32_t *cigar;
    const uint8_t *seq4;
    const int64_t *seq_off;
    const int2 *tile_rng;
    const uint8_t *ref;
    int64_t ref_start, ref_len;
    int64_t start, end;
    int32_t no_del;
    int32_t need_raw;      // read-level filters active: the row mask needs its own counter
    int32_t *out;
    uint8_t *rowmask;
    SmallState *st;
};

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------- small helpers

__device__ __forceinline__ int ref_code_of(int c) {
    if (c >= 'a' && c <= 'z') c -= 32;
    switch (c) {
    case '=': return 0; case 'A': return 1; case 'C': return 2; case 'M': return 3;
    case 'G': return 4; case 'R': return 5; case 'S': return 6; case 'V': return 7;
    case 'T': return 8; case 'W': return 9; case 'Y': return 10; case 'H': return 11;
    case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15;
    default: return 16;
    }
}

// 4-bit code -> mismatch class A,C,G,T,N,other (0..5), one nibble per code
#define MMCLASS_LUT 0x4555555355525105ull
__device__ __forceinline__ int mm_class(int code) { return (int)((MMCLASS_LUT >> (4 * code)) & 15ull); }

struct TileCtx {
    uint32_t *cnt;          // counter arrays, [array][TT]
    const uint32_t *seqw;   // staged sequence bytes viewed as words
    const uint32_t *refn;   // reference tile, 4-bit codes, first position in the top nibble
    const uint8_t *refc;    // reference code per position (0..16)
    int t0, t1;             // tile position range (t1 = t0 + TT)
    int no_del;
    bool slow;              // a reference position without a code: compare base by base
};

__device__ __forceinline__ void interval_add(uint32_t *arr, int x, int l, int t0) {
    int a = max(x - t0, 0), b = min(x + l - t0, TT);
    if (a < b) {
        atomicAdd(arr + a, 1u);
        if (b < TT) atomicAdd(arr + b, 0xffffffffu);
    }
}

// ---- variation: one aligned (M / = / X) segment ---------------------------------------
// x: reference start, l: length, y: query offset of its first base, sn0: nibble index of the
// read's first base in the staged sequence stream.
__device__ __forceinline__ void var_match_segment(const TileCtx &T, int bank, int x, int l, int y, uint32_t sn0,
                                                  bool hasseq, bool ins) {
    const int a = max(x - T.t0, 0), b = min(x + l - T.t0, TT);
    if (a >= b) return;
    uint32_t *nondel = T.cnt + bank * TT;
    atomicAdd(nondel + a, 1u);
    if (b < TT) atomicAdd(nondel + b, 0xffffffffu);
    uint32_t *pt = T.cnt + KindInfo<KIND_VAR>::NDIFF * TT;
    if (ins) {   // htslib indel > 0 at the last base of the op (A.1), opt.pyx:501
        int li = x + l - 1 - T.t0;
        if (li >= 0 && li < TT) atomicAdd(pt + bank * TT + li, 1u);
    }
    uint32_t *mm = pt + (4 + bank) * TT;       // class c at mm + 4 c TT
    const int len = b - a;
    if (!hasseq) {                              // no sequence: every base is an "other" mismatch (opt.pyx:2146-2147)
        for (int i = 0; i < len; ++i) atomicAdd(mm + 5 * 4 * TT + a + i, 1u);
        return;
    }
    uint32_t sn = sn0 + (uint32_t)(y + (a + T.t0 - x));
    if (T.slow) {
        for (int i = 0; i < len; ++i) {
            uint32_t k = sn + i;
            uint32_t w = T.seqw[k >> 3];
            int code = (w >> (((k & 7) ^ 1) << 2)) & 15;   // byte k/2 of the stream, high nibble first
            if (code != T.refc[a + i]) atomicAdd(mm + mm_class(code) * 4 * TT + a + i, 1u);
        }
        return;
    }
    uint32_t si = sn >> 3, ssh = (sn & 7) << 2;
    uint32_t ri = (uint32_t)a >> 3, rsh = ((uint32_t)a & 7) << 2;
    uint32_t s_hi = __byte_perm(T.seqw[si], 0, 0x0123), r_hi = T.refn[ri];
    for (int i = 0; i < len; i += 8) {
        uint32_t s_lo = __byte_perm(T.seqw[++si], 0, 0x0123), r_lo = T.refn[++ri];
        uint32_t sw = __funnelshift_l(s_lo, s_hi, ssh), rw = __funnelshift_l(r_lo, r_hi, rsh);
        s_hi = s_lo; r_hi = r_lo;
        uint32_t xr = sw ^ rw;
        int rem = len - i;
        if (rem < 8) xr &= 0xffffffffu << ((8 - rem) << 2);
        while (xr) {
            int lz = __clz(xr) >> 2;
            int sh = 28 - (lz << 2);
            int code = (sw >> sh) & 15;
            atomicAdd(mm + mm_class(code) * 4 * TT + a + i + lz, 1u);
            xr &= ~(0xfu << sh);
        }
    }
}

// ---- generic per-read walk ----------------------------------------------------------------

template <int KIND>
__device__ __forceinline__ void cover_segment(const TileCtx &T, uint32_t cls, int x, int l) {
    if (KIND == KIND_COV) interval_add(T.cnt + (cls & 3u) * TT, x, l, T.t0);
    else {   // KIND_EXT: one array per (class, strand)
        uint32_t m = (cls >> CLS_EXT_SHIFT) & 255u;
        const int s = cls & CLS_REV ? 1 : 0;
        const int a = max(x - T.t0, 0), b = min(x + l - T.t0, TT);
        if (a >= b) return;
        while (m) {
            int k = __ffs(m) - 1;
            m &= m - 1;
            uint32_t *arr = T.cnt + (k * 2 + s) * TT;
            atomicAdd(arr + a, 1u);
            if (b < TT) atomicAdd(arr + b, 0xffffffffu);
        }
    }
}

template <int KIND>
__device__ void tile_read(const TileCtx &T, const TileArgs &A, const RInfo &r, uint32_t sn0, uint32_t *raw) {
    const uint32_t cls = r.cls;
    if (!(cls & CLS_ADMITTED)) return;
    if (raw) interval_add(raw, r.pos, r.end - r.pos, T.t0);
    if (!(cls & CLS_COUNTED)) return;
    const int bank = cls & 3u;
    if (KIND != KIND_VAR && !T.no_del) {   // D / N / zero-length alignments all count (opt.pyx:1904-1906)
        cover_segment<KIND>(T, cls, r.pos, r.end - r.pos);
        return;
    }
    const bool hasseq = cls & CLS_HASSEQ;
    if (cls & CLS_SIMPLE) {
        if (KIND == KIND_VAR) var_match_segment(T, bank, r.pos, r.end - r.pos, 0, sn0, hasseq, false);
        else cover_segment<KIND>(T, cls, r.pos, r.end - r.pos);
        return;
    }
    if (cls & CLS_NOREF) {   // mapped, CIGAR consumes no reference: one deletion event at pos
        if (!T.no_del) {
            if (KIND == KIND_VAR) interval_add(T.cnt + (4 + bank) * TT, r.pos, 1, T.t0);
            else cover_segment<KIND>(T, cls, r.pos, 1);
        }
        return;
    }
    const uint32_t *cg = A.cigar + r.cigar_off;
    const int nc = (int)r.n_cigar;
    int x = r.pos, y = 0;
    for (int k = 0; k < nc && x < T.t1; ++k) {
        uint32_t w = __ldg(cg + k);
        int op = w & 15, l = (int)(w >> 4);
        if (op_is_match(op)) {
            if (x + l > T.t0) {
                if (KIND == KIND_VAR) {
                    bool ins = false;
                    if (k + 1 < nc && x + l - 1 >= T.t0 && x + l - 1 < T.t1) {
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
                    var_match_segment(T, bank, x, l, y, sn0, hasseq, ins);
                } else cover_segment<KIND>(T, cls, x, l);
            }
            x += l; y += l;
        } else if (op == OP_D || op == OP_N) {
            if (!T.no_del && x + l > T.t0) {
                if (KIND == KIND_VAR) interval_add(T.cnt + ((op == OP_N ? 8 : 4) + bank) * TT, x, l, T.t0);
                else cover_segment<KIND>(T, cls, x, l);
            }
            x += l;
        } else if (op == OP_I || op == OP_S) y += l;
    }
}

// ---- in-place inclusive prefix sum of one TT-word array by one warp --------------------------
// lane l owns words [128 k + 4 l, 128 k + 4 l + 4) for k = 0 .. TT/128-1 (conflict-free 128-bit
// accesses).
__device__ __forceinline__ void warp_scan_array(uint32_t *arr, int lane) {
    const int NK = TT / 128;
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

__device__ __forceinline__ void six(const uint32_t *v, int32_t *q, int64_t st) {
    // banks rev | pp << 1 -> all, fwd, rev, pp, pp_fwd, pp_rev (CountPpStrand, opt.pyx:147-168)
    uint32_t fwd = v[0] + v[2], rev = v[1] + v[3];
    q[0] = fwd + rev; q[st] = fwd; q[2 * st] = rev; q[3 * st] = v[2] + v[3]; q[4 * st] = v[2]; q[5 * st] = v[3];
}

template <int STAT>
__device__ __forceinline__ uint32_t tile_store(const uint32_t *cnt, int i, int refc, int32_t *o, int64_t st) {
    if (STAT == PSS_COVERAGE || STAT == PSS_COVERAGE_STRAND || STAT == PSS_COVERAGE_GC) {
        uint32_t n[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) n[b] = cnt[b * TT + i];
        uint32_t all = n[0] + n[1] + n[2] + n[3], pp = n[2] + n[3];
        if (STAT == PSS_COVERAGE) { o[0] = all; o[st] = pp; }
        else if (STAT == PSS_COVERAGE_GC) { o[st] = all; o[2 * st] = pp; }     // field 0 = gc (k_gc)
        else six(n, o, st);
        return all;
    } else if (STAT == PSS_COVERAGE_EXT || STAT == PSS_COVERAGE_EXT_STRAND) {
        uint32_t all = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t f = cnt[(2 * k) * TT + i], r = cnt[(2 * k + 1) * TT + i];
            if (k == 0) all = f + r;
            if (STAT == PSS_COVERAGE_EXT) o[k * st] = f + r;
            else { o[(3 * k) * st] = f + r; o[(3 * k + 1) * st] = f; o[(3 * k + 2) * st] = r; }
        }
        return all;
    } else {
        // groups: reads, matches, mismatches, deletions, insertions, A, C, T, G, N
        uint32_t g[10][4];
        const uint32_t *pt = cnt + KindInfo<KIND_VAR>::NDIFF * TT;
        uint32_t all = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            uint32_t nd = cnt[b * TT + i], dl = cnt[(4 + b) * TT + i], sk = cnt[(8 + b) * TT + i];
            uint32_t m[6], mm = 0;
#pragma unroll
            for (int c = 0; c < 6; ++c) { m[c] = pt[(4 + 4 * c + b) * TT + i]; mm += m[c]; }
            uint32_t mt = nd - mm;
            g[0][b] = nd + dl + sk; g[1][b] = mt; g[2][b] = mm; g[3][b] = dl; g[4][b] = pt[b * TT + i];
            g[5][b] = m[0] + (refc == 1 ? mt : 0u);    // A
            g[6][b] = m[1] + (refc == 2 ? mt : 0u);    // C
            g[7][b] = m[3] + (refc == 8 ? mt : 0u);    // T
            g[8][b] = m[2] + (refc == 4 ? mt : 0u);    // G
            g[9][b] = m[4] + (refc == 15 ? mt : 0u);   // N
            all += g[0][b];
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            if (STAT == PSS_VARIATION) {
                o[(2 * k) * st] = g[k][0] + g[k][1] + g[k][2] + g[k][3];
                o[(2 * k + 1) * st] = g[k][2] + g[k][3];
            } else six(g[k], o + (int64_t)(6 * k) * st, st);
        }
        return all;
    }
}

// ------------------------------------------------------------------------------ the kernel

template <int KIND, int STAT>
__global__ void __launch_bounds__(TTHREADS, 2)
k_tile(TileArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int NARR = KindInfo<KIND>::NDIFF + KindInfo<KIND>::NPOINT;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);
    uint32_t *raw = cnt + NARR * TT;                                   // only touched when need_raw
    RInfo *s_ri = reinterpret_cast<RInfo *>(raw + TT);
    // the rest exists for KIND_VAR only (see tile_smem_bytes)
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_ri + TCH);          // TSEQ_BYTES + 32
    uint32_t *s_refn = reinterpret_cast<uint32_t *>(s_seq + TSEQ_BYTES + 32);   // TT / 8 + 2 words
    uint8_t *s_refc = reinterpret_cast<uint8_t *>(s_refn + TT / 8 + 2);         // TT bytes
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_n, s_slow;
    __shared__ long long s_seqlo;      // global byte index staged at s_seq[0]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int64_t t0l = A.start + (int64_t)tile * TT;
    const int2 rng = A.tile_rng[tile];

    if (tid == 0) { mbar_init(&s_bar, 1); s_slow = 0; }
    {   // zero the counters
        const int nw4 = (NARR + (A.need_raw ? 1 : 0)) * TT / 4;
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < nw4; k += TTHREADS) reinterpret_cast<uint4 *>(cnt)[k] = z;
    }
    __syncthreads();
    if (KIND == KIND_VAR) {   // reference tile: code per position + packed nibbles
#pragma unroll
        for (int h = 0; h < TT / TTHREADS; ++h) {
            int i = tid + h * TTHREADS;
            int64_t ro = t0l + i - A.ref_start;
            int c = 16;
            if (A.ref && ro >= 0 && ro < A.ref_len) c = ref_code_of(__ldg(A.ref + ro));
            s_refc[i] = (uint8_t)c;
            if (c == 16 && t0l + i < A.end) s_slow = 1;
            uint32_t v = (uint32_t)(c & 15) << (28 - ((lane & 7) << 2));
            v |= __shfl_xor_sync(FULL, v, 1);
            v |= __shfl_xor_sync(FULL, v, 2);
            v |= __shfl_xor_sync(FULL, v, 4);
            if ((lane & 7) == 0) s_refn[i >> 3] = v;
        }
        if (tid < 2) s_refn[TT / 8 + tid] = 0;
    }

    TileCtx T;
    T.cnt = cnt; T.seqw = reinterpret_cast<const uint32_t *>(s_seq); T.refn = s_refn; T.refc = s_refc;
    T.t0 = (int)t0l; T.t1 = T.t0 + TT; T.no_del = A.no_del; T.slow = false;

    uint32_t parity = 0;
    for (int cb = rng.x; cb < rng.y;) {
        __syncthreads();                       // staging buffers free, s_slow / counters visible
        if (tid == 0) {
            int n = min(TCH, rng.y - cb);
            uint32_t bytes = 0;
            const uint8_t *src = nullptr;
            uint32_t sbytes = 0;
            if (KIND == KIND_VAR) {
                long long lo = A.seq_off[cb], hi = A.seq_off[cb + n];
                while (n > 1 && hi - lo + 32 > TSEQ_BYTES) { n >>= 1; hi = A.seq_off[cb + n]; }
                if (hi - lo + 32 > TSEQ_BYTES) { atomicOr(&A.st->toolong, 1); n = 0; }
                else {
                    uintptr_t a0 = reinterpret_cast<uintptr_t>(A.seq4 + lo), a1 = reinterpret_cast<uintptr_t>(A.seq4 + hi);
                    uintptr_t al = a0 & ~(uintptr_t)15, ah = (a1 + 15) & ~(uintptr_t)15;
                    src = reinterpret_cast<const uint8_t *>(al);
                    sbytes = (uint32_t)(ah - al);
                    s_seqlo = lo - (long long)(a0 - al);
                }
            }
            s_n = n;
            fence_proxy_async();
            if (n > 0) {
                bytes = (uint32_t)n * (uint32_t)sizeof(RInfo) + sbytes;
                mbar_arrive_expect_tx(&s_bar, bytes);
                bulk_g2s(s_ri, A.ri + cb, (uint32_t)n * (uint32_t)sizeof(RInfo), &s_bar);
                if (sbytes) bulk_g2s(s_seq, src, sbytes, &s_bar);
            } else mbar_arrive_expect_tx(&s_bar, 0);
        }
        mbar_wait(&s_bar, parity);
        parity ^= 1;
        const int n = s_n;
        if (n == 0) break;
        T.slow = s_slow != 0;
        if (tid < n) {
            const RInfo r = s_ri[tid];
            uint32_t sn0 = (KIND == KIND_VAR) ? (uint32_t)((r.seq_off - s_seqlo) << 1) : 0u;
            tile_read<KIND>(T, A, r, sn0, A.need_raw ? raw : nullptr);
        }
        cb += n;
    }
    __syncthreads();
    {   // prefix sums of the difference arrays (+ raw)
        const int nscan = KindInfo<KIND>::NDIFF + (A.need_raw ? 1 : 0);
        for (int a = warp; a < nscan; a += TTHREADS / 32) {
            uint32_t *arr = (a < KindInfo<KIND>::NDIFF) ? cnt + a * TT : raw;
            warp_scan_array(arr, lane);
        }
    }
    __syncthreads();
    const int64_t npos = A.end - A.start;
#pragma unroll
    for (int h = 0; h < TT / TTHREADS; ++h) {
        int i = tid + h * TTHREADS;
        int64_t pl = t0l + i;
        if (pl < A.end) {
            int refc = (KIND == KIND_VAR) ? s_refc[i] : 16;
            uint32_t all = tile_store<STAT>(cnt, i, refc, A.out + (pl - A.start), npos);
            uint32_t rw = A.need_raw ? raw[i] : all;
            A.rowmask[pl - A.start] = rw ? 1 : 0;
        }
    }
}

template <int KIND>
static size_t tile_smem_bytes(bool need_raw) {
    size_t narr = KindInfo<KIND>::NDIFF + KindInfo<KIND>::NPOINT + 1;   // raw slot always reserved (layout)
    (void)need_raw;
    size_t b = narr * TT * 4 + sizeof(RInfo) * TCH;
    if (KIND == KIND_VAR) b += TSEQ_BYTES + 32 + (TT / 8 + 2) * 4 + TT;
    return b;
}

template <int KIND, int STAT>
static cudaError_t launch_tile_one(const TileArgs &a, int64_t ntiles, cudaStream_t s) {
    size_t smem = tile_smem_bytes<KIND>(a.need_raw);
    static bool configured = false;     // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tile<KIND, STAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_tile<KIND, STAT><<<(unsigned)ntiles, TTHREADS, smem, s>>>(a);
    return cudaGetLastError();
}

bool tile_kernel_supports(int stat, int min_baseq) {
    if (min_baseq > 0) return false;
    switch (stat) {
    case PSS_COVERAGE: case PSS_COVERAGE_STRAND: case PSS_COVERAGE_GC: case PSS_COVERAGE_EXT:
    case PSS_COVERAGE_EXT_STRAND: case PSS_VARIATION: case PSS_VARIATION_STRAND: return true;
    default: return false;
    }
}

cudaError_t launch_tile(int stat, const PileupArgs &p, const int64_t *seq_off, SmallState *st, int need_raw,
                        int64_t ntiles, cudaStream_t s, int64_t *launches) {
    if (ntiles <= 0) return cudaSuccess;
    TileArgs a;
    a.ri = p.ri; a.cigar = p.cigar; a.seq4 = p.seq4; a.seq_off = seq_off; a.tile_rng = p.tile_rng;
    a.ref = p.ref; a.ref_start = p.ref_start; a.ref_len = p.ref_len; a.start = p.start; a.end = p.end;
    a.no_del = p.no_del; a.need_raw = need_raw; a.out = p.out; a.rowmask = p.rowmask; a.st = st;
    if (launches) ++*launches;
    switch (stat) {
    case PSS_COVERAGE: return launch_tile_one<KIND_COV, PSS_COVERAGE>(a, ntiles, s);
    case PSS_COVERAGE_STRAND: return launch_tile_one<KIND_COV, PSS_COVERAGE_STRAND>(a, ntiles, s);
    case PSS_COVERAGE_GC: return launch_tile_one<KIND_COV, PSS_COVERAGE_GC>(a, ntiles, s);
    case PSS_COVERAGE_EXT: return launch_tile_one<KIND_EXT, PSS_COVERAGE_EXT>(a, ntiles, s);
    case PSS_COVERAGE_EXT_STRAND: return launch_tile_one<KIND_EXT, PSS_COVERAGE_EXT_STRAND>(a, ntiles, s);
    case PSS_VARIATION: return launch_tile_one<KIND_VAR, PSS_VARIATION>(a, ntiles, s);
    case PSS_VARIATION_STRAND: return launch_tile_one<KIND_VAR, PSS_VARIATION_STRAND>(a, ntiles, s);
    default: return cudaErrorInvalidValue;
    }
}
