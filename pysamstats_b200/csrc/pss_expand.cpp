// Host side of the narrow result transport (see pss_pileup_host_multi): the counting statistics
// come back from the device as 8-bit or 16-bit columns (a quarter / half of the PCIe bytes and of
// the staging traffic in host memory, which is what bounds the host path) and are widened here into
// the caller's int32 columns by a small pool of worker threads, with non-temporal stores so the
// host memory traffic is one read of the narrow data and one write of the wide data.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "pss_expand.h"

struct pss_pool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::function<void()>> queue;
    std::vector<char> finished;            // per job of `queue`
    size_t next = 0, done = 0, submitted = 0;
    size_t low = 0;                        // every job with index < low has finished
    size_t base = 0;                       // jobs of earlier (recycled) queues: tickets stay monotonic
    bool stop = false;
};

pss_pool *pss_pool_create(int n_threads) {
    if (n_threads <= 0) {
        // PSS_HOST_THREADS, else the host's cores shared between the ranks of this node (torchrun
        // exports LOCAL_WORLD_SIZE), capped at 32
        const char *e = getenv("PSS_HOST_THREADS");
        if (e && atoi(e) > 0) n_threads = atoi(e);
        else {
            unsigned hc = std::thread::hardware_concurrency();
            n_threads = hc ? (int)hc : 4;
            const char *lw = getenv("LOCAL_WORLD_SIZE");
            if (lw && atoi(lw) > 1) n_threads = n_threads / atoi(lw);
            if (n_threads > 32) n_threads = 32;
            if (n_threads < 2) n_threads = 2;
        }
    }
    pss_pool *p = new pss_pool();
    for (int t = 0; t < n_threads; ++t)
        p->threads.emplace_back([p] {
            for (;;) {
                std::function<void()> job;
                size_t idx;
                {
                    std::unique_lock<std::mutex> lk(p->mu);
                    p->cv_work.wait(lk, [p] { return p->stop || p->next < p->queue.size(); });
                    if (p->next >= p->queue.size()) return;      // stop requested and nothing left
                    idx = p->next++;
                    job = std::move(p->queue[idx]);
                }
                job();
                {
                    std::lock_guard<std::mutex> lk(p->mu);
                    ++p->done;
                    p->finished[idx] = 1;
                    while (p->low < p->finished.size() && p->finished[p->low]) ++p->low;
                }
                p->cv_done.notify_all();
            }
        });
    return p;
}

void pss_pool_destroy(pss_pool *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_work.notify_all();
    for (auto &t : p->threads) t.join();
    delete p;
}

int pss_pool_size(const pss_pool *p) { return p ? (int)p->threads.size() : 0; }

// returns a ticket: pss_pool_wait(ticket) returns once every job submitted up to it has run
size_t pss_pool_submit(pss_pool *p, std::function<void()> job) {
    size_t ticket;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        if (p->next == p->queue.size() && p->done == p->submitted) {   // idle: recycle the queue storage
            p->base += p->queue.size();
            p->queue.clear();
            p->finished.clear();
            p->next = 0;
            p->done = p->submitted = 0;
            p->low = 0;
        }
        p->queue.push_back(std::move(job));
        p->finished.push_back(0);
        ++p->submitted;
        ticket = p->base + p->queue.size();
    }
    p->cv_work.notify_one();
    return ticket;
}

void pss_pool_wait(pss_pool *p, size_t ticket) {
    std::unique_lock<std::mutex> lk(p->mu);
    // jobs finish out of order: `low` is the watermark below which every job of the current queue
    // has finished; every job of an earlier queue has (the queue is only recycled when idle)
    p->cv_done.wait(lk, [&] { return p->base + p->low >= ticket; });
}

__attribute__((target("avx2"))) static void widen_avx2(const uint16_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        __m256i a = _mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i)));
        __m256i b = _mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 8)));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 8), b);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

__attribute__((target("avx2"))) static void widen8_avx2(const uint8_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_cvtepu8_epi32(v));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(v, 8)));
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

// AVX-512: one non-temporal store per 64-byte line of the destination -- a whole line leaves the write-combining buffer
// at once (measured with pss_host_bandwidth on the build host: 2-4x the bytes per second of 16- / 32-byte streaming stores)
__attribute__((target("avx512f,avx512bw"))) static void widen8_avx512(const uint8_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16)
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), _mm512_cvtepu8_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
__attribute__((target("avx512f,avx512bw"))) static void widen16_avx512(const uint16_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16)
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), _mm512_cvtepu16_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

static int widen_isa() {          // 2: AVX-512, 1: AVX2, 0: scalar (PSS_WIDEN_ISA overrides downwards: measurements)
    static const int isa = [] {
        int v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
        if (const char *e = getenv("PSS_WIDEN_ISA")) v = std::min(v, atoi(e));
        return v;
    }();
    return isa;
}

void pss_widen_u8(const uint8_t *src, int32_t *dst, size_t n) {
    const int isa = widen_isa();
    if (isa == 2) { widen8_avx512(src, dst, n); return; }
    if (isa == 1) { widen8_avx2(src, dst, n); return; }
    for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}

void pss_widen_u16(const uint16_t *src, int32_t *dst, size_t n) {
    const int isa = widen_isa();
    if (isa == 2) { widen16_avx512(src, dst, n); return; }
    if (isa == 1) { widen_avx2(src, dst, n); return; }
    for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}

// ------------------------------------------------------------------------------ row sink

int64_t pss_count_rows(const uint8_t *mask, int64_t n) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) c += mask[i] != 0;
    return c;
}

template <class T>
static inline void pack_field(const T *col, const int32_t *sel, int m, uint8_t *o, int32_t row_bytes, int fb) {
    if (fb == 4) {
        for (int r = 0; r < m; ++r, o += row_bytes) { int32_t v = (int32_t)col[sel[r]]; memcpy(o, &v, 4); }
    } else {
        for (int r = 0; r < m; ++r, o += row_bytes) *o = (uint8_t)col[sel[r]];
    }
}

// 16 positions x 16 one-byte fields -> 16 records: the byte columns are transposed in registers (four rounds of
// unpacks) and every record receives its 16 fields as one widened 64-byte run.  col(f) = column f at the first
// position; `nst` leading int32 of the run are stored (16, or 8 for an eight-field statistic).
__attribute__((target("avx2"))) static inline void pack16x16_u8(const uint8_t *c0, size_t col_stride, int ncols, uint8_t *out,
                                                                 int32_t row_bytes) {
    __m128i a[16], t[16];
    for (int f = 0; f < 16; ++f)
        a[f] = f < ncols ? _mm_loadu_si128(reinterpret_cast<const __m128i *>(c0 + (size_t)f * col_stride)) : _mm_setzero_si128();
    for (int k = 0; k < 8; ++k) {                 // bytes of field pairs
        t[2 * k] = _mm_unpacklo_epi8(a[2 * k], a[2 * k + 1]);
        t[2 * k + 1] = _mm_unpackhi_epi8(a[2 * k], a[2 * k + 1]);
    }
    for (int q = 0; q < 4; ++q) {                 // fields 4q .. 4q + 3, positions 0-3 | 4-7 | 8-11 | 12-15
        a[4 * q] = _mm_unpacklo_epi16(t[4 * q], t[4 * q + 2]);
        a[4 * q + 1] = _mm_unpackhi_epi16(t[4 * q], t[4 * q + 2]);
        a[4 * q + 2] = _mm_unpacklo_epi16(t[4 * q + 1], t[4 * q + 3]);
        a[4 * q + 3] = _mm_unpackhi_epi16(t[4 * q + 1], t[4 * q + 3]);
    }
    for (int h = 0; h < 2; ++h)                   // fields 8h .. 8h + 7, two positions per register
        for (int j = 0; j < 4; ++j) {
            t[8 * h + 2 * j] = _mm_unpacklo_epi32(a[8 * h + j], a[8 * h + 4 + j]);
            t[8 * h + 2 * j + 1] = _mm_unpackhi_epi32(a[8 * h + j], a[8 * h + 4 + j]);
        }
    for (int k = 0; k < 8; ++k) {                 // all 16 fields of positions 2k, 2k + 1
        __m128i p0 = _mm_unpacklo_epi64(t[k], t[8 + k]), p1 = _mm_unpackhi_epi64(t[k], t[8 + k]);
        uint8_t *o0 = out + (size_t)(2 * k) * row_bytes, *o1 = o0 + row_bytes;
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(o0), _mm256_cvtepu8_epi32(p0));
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(o1), _mm256_cvtepu8_epi32(p1));
        if (ncols > 8) {
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(o0 + 32), _mm256_cvtepu8_epi32(_mm_srli_si128(p0, 8)));
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(o1 + 32), _mm256_cvtepu8_epi32(_mm_srli_si128(p1, 8)));
        }
    }
}

// the value fields of 64 consecutive positions (all emitted) from one-byte columns, fields consecutive in the record
__attribute__((target("avx2"))) static void pack_dense_u8(const pss_row_pack *K, const uint8_t *src, size_t stride, int64_t i0,
                                                           int m, uint8_t *out) {
    const int nw = K->n_want;
    const uint8_t *c0 = src + (size_t)K->f_idx[0] * stride + i0;
    uint8_t *o = out + K->f_off[0];
    for (int w = 0; w < nw; ++w) __builtin_prefetch(c0 + (size_t)w * stride + 128);
    for (int p = 0; p + 16 <= m; p += 16) {
        if (nw == 8) pack16x16_u8(c0 + p, stride, 8, o + (size_t)p * K->row_bytes, K->row_bytes);
        else
            for (int g = 0; g < nw; g += 16) {
                const int gg = g + 16 <= nw ? g : nw - 16;      // the last group overlaps the one before
                pack16x16_u8(c0 + (size_t)gg * stride + p, stride, 16, o + (size_t)p * K->row_bytes + 4 * gg, K->row_bytes);
            }
    }
}

// block of finished records -> its place in the caller's array with non-temporal stores: the records are never
// read back by the workers, and a regular store would first fetch every destination line from memory
__attribute__((target("avx512f"))) static void copy_out_nt512(uint8_t *dst, const uint8_t *src, size_t n) {
    for (; n >= 64; n -= 64, dst += 64, src += 64)       // one store per destination line
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst), _mm512_loadu_si512(reinterpret_cast<const void *>(src)));
    memcpy(dst, src, n);
}
__attribute__((target("avx2"))) static void copy_out_nt(uint8_t *dst, const uint8_t *src, size_t n) {
    size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    if (widen_isa() == 2) { copy_out_nt512(dst, src, n); return; }
    for (; n >= 64; n -= 64, dst += 64, src += 64) {
        __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
        __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 32), b);
    }
    memcpy(dst, src, n);
}

static bool pack_simd_ok(const pss_row_pack *K) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (!avx2 || !(K->n_want == 8 || K->n_want >= 16)) return false;
    for (int w = 0; w < K->n_want; ++w)
        if (K->f_bytes[w] != 4 || K->f_idx[w] != K->f_idx[0] + w || K->f_off[w] != K->f_off[0] + 4 * w) return false;
    // the 64-byte runs of a record's last group may not reach past the record (they are overwritten by nothing)
    return K->f_off[0] + 4 * K->n_want <= K->row_bytes;
}

void pss_pack_piece(pss_row_pack *K, const uint8_t *src, int64_t stride, int esz, const uint8_t *mask, int64_t p0,
                    int64_t len, int64_t row0) {
    const int BLK = 64;               // 64 records of ~250 bytes stay in L1 while the columns stream through
    int32_t sel[BLK + 1];
    const int32_t rb = K->row_bytes;
    const bool simd = esz == 1 && pack_simd_ok(K);
    static const bool avx2 = __builtin_cpu_supports("avx2");
    static thread_local std::vector<uint8_t> block;      // the records of a block are assembled here (L1) ...
    if (avx2) block.assign((size_t)BLK * (size_t)rb + 64, 0);     // (bytes of a record that no field covers leave as zeros)
    int64_t i = 0, row = row0, neg = 0;
    while (i < len) {
        int m = 0;
        const int64_t i0 = i;
        while (i < len && m < BLK) { sel[m] = (int32_t)i; m += mask[i] != 0; ++i; }
        if (m == 0) break;
        uint8_t *const dst = K->rows + (size_t)row * (size_t)rb;
        uint8_t *out = avx2 ? block.data() : dst;
        const bool dense = simd && m == BLK && i - i0 == BLK;       // 64 consecutive positions, all emitted
        if (dense) pack_dense_u8(K, src, (size_t)stride, i0, m, out);
        for (int w = 0; w < (dense ? 0 : K->n_want); ++w) {
            const uint8_t *col = src + (size_t)K->f_idx[w] * (size_t)stride * (size_t)esz;
            uint8_t *o = out + K->f_off[w];
            if (esz == 1) pack_field((const uint8_t *)col, sel, m, o, rb, K->f_bytes[w]);
            else if (esz == 2) pack_field((const uint16_t *)col, sel, m, o, rb, K->f_bytes[w]);
            else {
                const int32_t *c4 = (const int32_t *)col;
                if (K->f_idx[w] == K->check_field)
                    for (int r = 0; r < m; ++r) neg += c4[sel[r]] < 0;
                pack_field(c4, sel, m, o, rb, K->f_bytes[w]);
            }
        }
        if (K->chrom_off >= 0) {
            uint8_t *o = out + K->chrom_off;
            for (int r = 0; r < m; ++r, o += rb) memcpy(o, K->chrom, (size_t)K->chrom_bytes);
        }
        if (K->pos_off >= 0) {
            uint8_t *o = out + K->pos_off;
            for (int r = 0; r < m; ++r, o += rb) { int32_t v = (int32_t)(p0 + sel[r] + K->pos_add); memcpy(o, &v, 4); }
        }
        if (K->ref_off >= 0) {
            uint8_t *o = out + K->ref_off;
            for (int r = 0; r < m; ++r, o += rb) {
                int64_t ro = p0 + sel[r] - K->ref_start;
                uint8_t c = (ro >= 0 && ro < K->ref_len) ? K->ref[ro] : 0;
                if (c >= 'a' && c <= 'z') c -= 32;
                *o = c;
            }
        }
        if (avx2) copy_out_nt(dst, out, (size_t)m * (size_t)rb);     // ... and leave with streaming stores
        row += m;
    }
    if (avx2) _mm_sfence();
    if (neg) __atomic_fetch_add(&K->n_negative, neg, __ATOMIC_RELAXED);
}
