// Host-side BAM decoder / writer behind include/pss_host.h: multi-threaded BGZF inflate,
// BAM record indexing and parallel structure-of-arrays fill.  Plain C++17 + zlib + pthreads.
//
// BGZF (SAM spec section 4.1): a series of gzip members, each with a "BC" extra subfield
// holding the block size, each inflating to <= 64 KiB, independently decodable -- so blocks
// are distributed over threads.  BAM (section 4.2): after the header, records
// {block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID,
//  next_pos, tlen, read_name, cigar, seq, qual, aux}.
#include "../../include/pss_host.h"
#include <immintrin.h>

#include "pss_inflate.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *pss_host_error(void) { return g_err; }

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int32_t rdi32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }

static int inflate_one(z_stream &zs, const uint8_t *src, int64_t in_len, uint8_t *dst, int64_t out_len);

template <class F>
static void parallel_for(int64_t n, int n_threads, F f) {
    if (n_threads <= 1 || n < 2) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    int64_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=] { f(lo, hi, t); });
    }
    for (auto &x : th) x.join();
}

// Large byte buffer that is NOT zero-filled on allocation and asks for transparent huge pages: the
// inflate threads are the first to touch it, and with 4 KiB pages their page faults (serialised in
// the kernel) cost more than the inflation itself.
struct BigBuf {
    uint8_t *p = nullptr;
    size_t n = 0, cap = 0;
    BigBuf() {}
    BigBuf(const BigBuf &) = delete;
    BigBuf &operator=(const BigBuf &) = delete;
    ~BigBuf() { release(); }
    void release() { if (p) munmap(p, cap); p = nullptr; n = cap = 0; }
    bool resize(size_t bytes) {         // contents are not preserved
        release();
        if (bytes == 0) { n = 0; return true; }
        const size_t huge = (size_t)2 << 20;
        cap = (bytes + huge - 1) / huge * huge;
        void *m = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) { cap = 0; return false; }
#ifdef MADV_HUGEPAGE
        madvise(m, cap, MADV_HUGEPAGE);
#endif
        p = (uint8_t *)m; n = bytes;
        return true;
    }
    uint8_t *data() { return p; }
    const uint8_t *data() const { return p; }
    size_t size() const { return n; }
};

// read-only mapping of a whole file (no copy into a buffer of our own)
struct FileMap {
    const uint8_t *p = nullptr;
    size_t n = 0;
    ~FileMap() { if (p) munmap((void *)p, n); }
    bool open(const char *path, bool whole = true) {      // whole: the entire file is about to be read front to back
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); return false; }
        n = (size_t)st.st_size;
        if (n == 0) { ::close(fd); p = nullptr; return true; }
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (m == MAP_FAILED) { n = 0; return false; }
        if (whole) {
            madvise(m, n, MADV_SEQUENTIAL);
            madvise(m, n, MADV_WILLNEED);
        }
        p = (const uint8_t *)m;
        return true;
    }
};

struct Blk { int64_t in_off, in_len, out_off, out_len, file_off; };   // payload span, inflated span, block start in the file

struct pss_bam {
    std::vector<Blk> blocks;             // BGZF blocks behind `raw` (whole-file open only)
    BigBuf raw;                          // inflated BAM stream
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    std::vector<int64_t> rec;            // offset of each record's refID field in raw
    std::vector<int64_t> cig_off, seq_off, qual_off;   // n + 1 prefix sums
    int32_t name_width = 1;
    int n_threads = 1;
    double t_read = 0, t_inflate = 0, t_index = 0;
};

static const int REF_CONSUMING[16] = {1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0};   // M I D N S H P = X

extern "C" int pss_bam_open(const char *path, int n_threads, pss_bam **out) {
    if (!out || !path) return fail(-1, "NULL argument");
    *out = nullptr;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    double t0 = now();
    FileMap fm;
    if (!fm.open(path)) return fail(-2, "cannot open %s", path);
    const long long fsize = (long long)fm.n;
    const uint8_t *comp = fm.p;
    double t1 = now();
    // block table
    std::vector<Blk> blocks;
    int64_t off = 0, total = 0;
    while (off < fsize) {
        if (off + 18 > fsize || comp[off] != 0x1f || comp[off + 1] != 0x8b || comp[off + 2] != 8 || !(comp[off + 3] & 4))
            return fail(-3, "not a BGZF file (bad block header at %lld)", (long long)off);
        int xlen = rd16(&comp[off + 10]);
        int64_t x = off + 12, xend = x + xlen;
        int bsize = -1;
        while (x + 4 <= xend) {
            int slen = rd16(&comp[x + 2]);
            if (comp[x] == 66 && comp[x + 1] == 67 && slen == 2) bsize = rd16(&comp[x + 4]);
            x += 4 + slen;
        }
        if (bsize < 0 || off + bsize + 1 > fsize) return fail(-3, "BGZF block without BC subfield / truncated at %lld", (long long)off);
        int64_t isize = rd32(&comp[off + bsize + 1 - 4]);
        blocks.push_back({xend, off + bsize + 1 - 8 - xend, total, isize, off});
        total += isize;
        off += bsize + 1;
    }
    pss_bam *b = new pss_bam();
    b->n_threads = n_threads;
    if (!b->raw.resize((size_t)total)) { delete b; return fail(-8, "cannot allocate %lld bytes", (long long)total); }
    std::atomic<int> bad(0);
    std::atomic<int64_t> next(0);
    // the blocks are inflated by n_threads workers while THIS thread already parses the header and walks the
    // block_size chain of the records behind the frontier of contiguously inflated blocks
    std::vector<std::atomic<uint8_t>> done(blocks.size());
    for (auto &d : done) d.store(0, std::memory_order_relaxed);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&] {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
            for (;;) {
                int64_t k0 = next.fetch_add(16);
                if (k0 >= (int64_t)blocks.size()) break;
                int64_t k1 = std::min<int64_t>(blocks.size(), k0 + 16);
                for (int64_t k = k0; k < k1; ++k) {
                    const Blk &bl = blocks[k];
                    if (inflate_one(zs, comp + bl.in_off, bl.in_len, b->raw.data() + bl.out_off, bl.out_len) != 0) bad = 1;
                    done[(size_t)k].store(1, std::memory_order_release);
                }
            }
            inflateEnd(&zs);
        });
    auto join_all = [&] { for (auto &x : th) if (x.joinable()) x.join(); };
    struct Joiner { std::vector<std::thread> &t; ~Joiner() { for (auto &x : t) if (x.joinable()) x.join(); } } joiner{th};
    size_t frontier = 0;
    int64_t avail = 0;
    auto wait_for = [&](int64_t need) -> bool {        // bytes [0, need) of the stream are inflated (need <= total)
        while (avail < need) {
            while (frontier < blocks.size() && done[frontier].load(std::memory_order_acquire)) ++frontier;
            avail = frontier < blocks.size() ? blocks[frontier].out_off : total;
            if (avail >= need) break;
            if (bad) return false;
            std::this_thread::yield();
        }
        return true;
    };
#define NEED(x) do { if (!wait_for(x)) { join_all(); delete b; return fail(-3, "corrupt BGZF block in %s", path); } } while (0)
    const uint8_t *r = b->raw.data();
    const int64_t n = total;
    if (n < 12) { join_all(); delete b; return fail(-4, "not a BAM file: %s", path); }
    NEED(12);
    if (memcmp(r, "BAM\1", 4) != 0) { join_all(); delete b; return fail(-4, "not a BAM file: %s", path); }
    int64_t p = 8 + rdi32(r + 4);
    if (p + 4 > n) { join_all(); delete b; return fail(-4, "truncated BAM header"); }
    NEED(p + 4);
    int32_t n_ref = rdi32(r + p);
    p += 4;
    for (int32_t i = 0; i < n_ref; ++i) {
        if (p + 4 > n) { join_all(); delete b; return fail(-4, "truncated BAM header"); }
        NEED(p + 4);
        int32_t l_name = rdi32(r + p);
        if (l_name < 1 || p + 8 + l_name > n) { join_all(); delete b; return fail(-4, "truncated BAM header"); }
        NEED(p + 8 + l_name);
        b->ref_names.emplace_back((const char *)(r + p + 4), (size_t)(l_name - 1));
        b->ref_lens.push_back(rdi32(r + p + 4 + l_name));
        p += 8 + l_name;
    }
    // record index + prefix sums (sequential: one pass over the block_size chain)
    int maxname = 1;
    b->cig_off.push_back(0); b->seq_off.push_back(0); b->qual_off.push_back(0);
    while (p < n) {
        if (p + 36 > n) { join_all(); delete b; return fail(-4, "truncated BAM record at %lld", (long long)p); }
        NEED(p + 36);
        int32_t bs = rdi32(r + p);
        if (bs < 32 || p + 4 + bs > n) { join_all(); delete b; return fail(-4, "bad BAM record size at %lld", (long long)p); }
        const uint8_t *q = r + p + 4;
        int l_name = q[8], n_cig = rd16(q + 12);
        int32_t l_seq = rdi32(q + 16);
        if (l_seq < 0 || 32 + (int64_t)l_name + 4ll * n_cig + (l_seq + 1) / 2 + l_seq > bs) {
            join_all();
            delete b;
            return fail(-4, "inconsistent BAM record at %lld", (long long)p);
        }
        b->rec.push_back(p + 4);
        b->cig_off.push_back(b->cig_off.back() + n_cig);
        b->seq_off.push_back(b->seq_off.back() + (l_seq + 1) / 2);
        b->qual_off.push_back(b->qual_off.back() + l_seq);
        maxname = std::max(maxname, l_name > 0 ? l_name - 1 : 0);
        p += 4 + bs;
    }
    NEED(total);
#undef NEED
    join_all();
    if (bad) { delete b; return fail(-3, "corrupt BGZF block in %s", path); }
    double t2 = now();          // inflate and the (overlapped) record index are done
    b->name_width = std::max(1, maxname);
    b->blocks.swap(blocks);
    b->t_read = t1 - t0; b->t_inflate = t2 - t1; b->t_index = now() - t2;
    *out = b;
    return 0;
}

static inline void wr16(std::vector<uint8_t> &v, uint16_t x) { v.push_back(x & 255); v.push_back(x >> 8); }
static inline void wr32(std::vector<uint8_t> &v, uint32_t x) { for (int k = 0; k < 4; ++k) v.push_back((x >> (8 * k)) & 255); }

static int reg2bin(int64_t beg, int64_t end) {     // SAM spec section 5.3
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// ------------------------------------------------------------------- BAI: region open + index build
//
// BAI (SAM spec section 5.2): per reference a binning index (bin -> chunks of virtual offsets
// [beg, end), virtual offset = block start in the file << 16 | offset in the inflated block) and a
// linear index (smallest virtual offset of a record overlapping each 16 kb window).

struct Chunk { uint64_t beg, end; };

static void reg2bins(int64_t beg, int64_t end, std::vector<uint32_t> &bins) {     // SAM spec section 5.3
    --end;
    bins.push_back(0);
    for (int64_t k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) bins.push_back((uint32_t)k);
    for (int64_t k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) bins.push_back((uint32_t)k);
    for (int64_t k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) bins.push_back((uint32_t)k);
    for (int64_t k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) bins.push_back((uint32_t)k);
    for (int64_t k = 4681 + (beg >> 14); k <= 4681 + (end >> 14); ++k) bins.push_back((uint32_t)k);
}

static bool read_file(const char *path, std::vector<uint8_t> &out) {
    FILE *fh = fopen(path, "rb");
    if (!fh) return false;
    fseek(fh, 0, SEEK_END);
    long long n = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    out.resize((size_t)n);
    bool ok = n == 0 || fread(out.data(), 1, (size_t)n, fh) == (size_t)n;
    fclose(fh);
    return ok;
}

// chunks of reference `tid` that can hold records overlapping [beg, end), sorted and merged
// (an index that cannot belong to this file -- other number of references, offsets beyond its end: a stale .bai --
// is refused with -6 and the caller decodes the whole file instead)
static int bai_chunks(const char *bai_path, int32_t tid, int64_t beg, int64_t end, std::vector<Chunk> &out,
                      int32_t bam_n_ref = -1, int64_t bam_size = -1) {
    std::vector<uint8_t> d;
    if (!read_file(bai_path, d)) return fail(-2, "cannot read index %s", bai_path);
    const int64_t n = (int64_t)d.size();
    if (n < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return fail(-6, "not a BAI index: %s", bai_path);
    const int32_t n_ref = rdi32(&d[4]);
    if (tid < 0 || tid >= n_ref) return fail(-6, "reference %d is not in the index (%d references)", tid, n_ref);
    if (bam_n_ref >= 0 && n_ref != bam_n_ref)
        return fail(-6, "index %s has %d references, the BAM header %d: stale index", bai_path, n_ref, bam_n_ref);
    std::vector<uint32_t> want;
    reg2bins(beg, end, want);
    std::sort(want.begin(), want.end());
    int64_t p = 8;
    auto rd64 = [&](int64_t o) { uint64_t v; memcpy(&v, &d[o], 8); return v; };
    for (int32_t r = 0; r <= tid; ++r) {
        if (p + 4 > n) return fail(-6, "truncated BAI index");
        const int32_t n_bin = rdi32(&d[p]);
        p += 4;
        for (int32_t k = 0; k < n_bin; ++k) {
            if (p + 8 > n) return fail(-6, "truncated BAI index");
            const uint32_t bin = rd32(&d[p]);
            const int32_t n_chunk = rdi32(&d[p + 4]);
            p += 8;
            if (n_chunk < 0 || p + 16ll * n_chunk > n) return fail(-6, "truncated BAI index");
            if (r == tid && bin != 37450u && std::binary_search(want.begin(), want.end(), bin))
                for (int32_t c = 0; c < n_chunk; ++c) out.push_back({rd64(p + 16ll * c), rd64(p + 16ll * c + 8)});
            p += 16ll * n_chunk;
        }
        if (p + 4 > n) return fail(-6, "truncated BAI index");
        const int32_t n_intv = rdi32(&d[p]);
        p += 4;
        if (n_intv < 0 || p + 8ll * n_intv > n) return fail(-6, "truncated BAI index");
        if (r == tid && n_intv > 0) {
            // every record overlapping the region overlaps window beg >> 14 or starts later:
            // its virtual offset is >= the linear index entry of that window
            int64_t w = beg >> 14;
            if (w >= n_intv) w = n_intv - 1;
            const uint64_t min_off = rd64(p + 8 * w);
            std::vector<Chunk> kept;
            for (Chunk c : out) {
                if (c.end <= min_off) continue;
                if (c.beg < min_off) c.beg = min_off;
                kept.push_back(c);
            }
            out.swap(kept);
        }
        p += 8ll * n_intv;
    }
    if (bam_size >= 0)
        for (const Chunk &c : out)
            if ((int64_t)(c.beg >> 16) > bam_size || (int64_t)(c.end >> 16) > bam_size)
                return fail(-6, "index %s points beyond the end of the BAM file: stale index", bai_path);
    std::sort(out.begin(), out.end(), [](const Chunk &a, const Chunk &b) { return a.beg < b.beg; });
    // merge chunks that overlap or sit within 64 KiB of file of each other: records read on the
    // way are filtered by position anyway, and no BGZF block is inflated twice
    std::vector<Chunk> merged;
    for (const Chunk &c : out) {
        if (!merged.empty() && (c.beg >> 16) <= (merged.back().end >> 16) + 65536) {
            if (c.end > merged.back().end) merged.back().end = c.end;
        } else merged.push_back(c);
    }
    out.swap(merged);
    return 0;
}

// BGZF block at file offset `off`: total size (0 = not a block header)
static int64_t bgzf_block_size(const uint8_t *h, int64_t avail, int64_t *payload_off) {
    if (avail < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
    const int xlen = rd16(h + 10);
    int64_t x = 12, xend = 12 + xlen;
    if (xend > avail) return 0;
    int bsize = -1;
    while (x + 4 <= xend) {
        const int slen = rd16(h + x + 2);
        if (h[x] == 66 && h[x + 1] == 67 && slen == 2) bsize = rd16(h + x + 4);
        x += 4 + slen;
    }
    if (bsize < 0) return 0;
    *payload_off = xend;
    return (int64_t)bsize + 1;
}

static int parse_header(pss_bam *b, const uint8_t *r, int64_t n, int64_t *consumed) {   // 1: incomplete
    if (n < 12) return 1;
    if (memcmp(r, "BAM\1", 4) != 0) return fail(-4, "not a BAM file");
    int64_t p = 8 + (int64_t)rdi32(r + 4);
    if (p + 4 > n) return 1;
    const int32_t n_ref = rdi32(r + p);
    p += 4;
    b->ref_names.clear(); b->ref_lens.clear();
    for (int32_t i = 0; i < n_ref; ++i) {
        if (p + 4 > n) return 1;
        const int32_t l_name = rdi32(r + p);
        if (l_name < 1) return fail(-4, "bad BAM header");
        if (p + 8 + l_name > n) return 1;
        b->ref_names.emplace_back((const char *)(r + p + 4), (size_t)(l_name - 1));
        b->ref_lens.push_back(rdi32(r + p + 4 + l_name));
        p += 8 + l_name;
    }
    *consumed = p;
    return 0;
}

static int inflate_one(z_stream &zs, const uint8_t *src, int64_t in_len, uint8_t *dst, int64_t out_len) {
    if (out_len == 0) return 0;
    // own decoder first (pss_inflate.cpp, ~2x zlib); whatever it declines is repeated with zlib, which decides
    static const bool zlib_only = getenv("PSS_ZLIB_INFLATE") != nullptr;
    if (!zlib_only && pss_inflate_fast(src, (size_t)in_len, dst, (size_t)out_len) == 0) return 0;
    inflateReset(&zs);
    zs.next_in = (Bytef *)src; zs.avail_in = (uInt)in_len;
    zs.next_out = dst; zs.avail_out = (uInt)out_len;
    const int rc = inflate(&zs, Z_FINISH);
    return (rc == Z_STREAM_END && zs.avail_out == 0) ? 0 : -1;
}

extern "C" int pss_bam_open_region(const char *path, const char *bai_path, int32_t tid, int64_t beg, int64_t end,
                                   int n_threads, pss_bam **out) {
    if (!out || !path) return fail(-1, "NULL argument");
    *out = nullptr;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const double t0 = now();
    FILE *fh = fopen(path, "rb");
    if (!fh) return fail(-2, "cannot open %s", path);
    fseek(fh, 0, SEEK_END);
    const int64_t fsize = ftell(fh);
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{fh};
    auto read_at = [&](int64_t off, int64_t len, uint8_t *dst) -> bool {
        if (off < 0 || off + len > fsize) return false;
        if (fseek(fh, (long)off, SEEK_SET) != 0) return false;
        return len == 0 || fread(dst, 1, (size_t)len, fh) == (size_t)len;
    };
    pss_bam *b = new pss_bam();
    b->n_threads = n_threads;
    struct Guard { pss_bam *b; ~Guard() { delete b; } } guard{b};
    // ---- header: inflate blocks from the start of the file until it parses
    {
        std::vector<uint8_t> head, comp(65536 + 64);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) return fail(-3, "zlib init failed");
        int64_t off = 0, consumed = 0;
        int rc = 1;
        while (rc == 1 && off < fsize) {
            const int64_t avail = std::min<int64_t>(fsize - off, (int64_t)comp.size());
            if (!read_at(off, avail, comp.data())) { rc = fail(-2, "read error on %s", path); break; }
            int64_t pay = 0;
            const int64_t bs = bgzf_block_size(comp.data(), avail, &pay);
            if (bs <= 0 || bs > avail) { rc = fail(-3, "not a BGZF file: %s", path); break; }
            const int64_t isize = rd32(&comp[bs - 4]);
            const size_t old = head.size();
            head.resize(old + (size_t)isize);
            if (inflate_one(zs, comp.data() + pay, bs - 8 - pay, head.data() + old, isize) != 0) { rc = fail(-3, "corrupt BGZF block in %s", path); break; }
            off += bs;
            rc = parse_header(b, head.data(), (int64_t)head.size(), &consumed);
        }
        inflateEnd(&zs);
        if (rc == 1) return fail(-4, "truncated BAM header in %s", path);
        if (rc != 0) return rc;
    }
    b->cig_off.push_back(0); b->seq_off.push_back(0); b->qual_off.push_back(0);
    if (tid < 0) {            // header only
        b->t_read = now() - t0;
        guard.b = nullptr;
        *out = b;
        return 0;
    }
    if (tid >= (int32_t)b->ref_names.size()) return fail(-1, "reference %d out of range", tid);
    std::string bai = bai_path ? bai_path : std::string(path) + ".bai";
    std::vector<Chunk> chunks;
    // a negative start (one_based=True with start=0 gives -1, opt.pyx:2116) selects what start 0 selects; the
    // linear index has no window -1 and reg2bins no bins below 0
    if (beg < 0) beg = 0;
    if (end < beg + 1) end = beg + 1;
    int rc = bai_chunks(bai.c_str(), tid, beg, end, chunks, (int32_t)b->ref_names.size(), fsize);
    if (rc != 0) return rc;
    // ---- the compressed bytes of every chunk, its block table
    struct Piece { int64_t file_lo, file_hi; size_t comp_off; int64_t first_blk, n_blk; int64_t start, stop; };   // start / stop: offsets in raw
    std::vector<Piece> pieces;
    // the compressed bytes are used where they lie in the page cache (a mapping of the file, no copy): comp_off and
    // the blocks' in_off are file offsets
    FileMap fmap;
    if (!fmap.open(path, false)) return fail(-2, "cannot map %s", path);
    struct { const uint8_t *p; const uint8_t *data() const { return p; } } comp{fmap.p};
    std::vector<Blk> blocks;
    int64_t total = 0;
    for (const Chunk &c : chunks) {
        const int64_t lo = (int64_t)(c.beg >> 16), last = (int64_t)(c.end >> 16);
        if (lo >= fsize) continue;
        // the block at `last` is needed up to c.end & 0xffff: read its header to learn its size
        int64_t hi = last;
        if ((c.end & 0xffff) != 0 && last < fsize) {
            uint8_t h[64];
            const int64_t avail = std::min<int64_t>(64, fsize - last);
            int64_t pay = 0;
            if (!read_at(last, avail, h)) return fail(-2, "read error on %s", path);
            const int64_t bs = bgzf_block_size(h, avail, &pay);
            if (bs <= 0) return fail(-6, "index %s does not match %s (no BGZF block at %lld)", bai.c_str(), path, (long long)last);
            hi = last + bs;
        }
        if (hi > fsize) hi = fsize;
        Piece pc{lo, hi, (size_t)lo, (int64_t)blocks.size(), 0, 0, 0};
        madvise((void *)(fmap.p + (lo & ~(int64_t)4095)), (size_t)(hi - (lo & ~(int64_t)4095)), MADV_WILLNEED);
        int64_t off = 0;
        pc.start = total + (int64_t)(c.beg & 0xffff);
        pc.stop = -1;
        while (off < hi - lo) {
            int64_t pay = 0;
            const int64_t bs = bgzf_block_size(comp.data() + pc.comp_off + off, hi - lo - off, &pay);
            if (bs <= 0 || off + bs > hi - lo)
                return fail(-6, "index %s does not match %s (bad BGZF block at %lld)", bai.c_str(), path, (long long)(lo + off));
            const int64_t isize = rd32(comp.data() + pc.comp_off + off + bs - 4);
            if (lo + off == last) pc.stop = total + (int64_t)(c.end & 0xffff);
            blocks.push_back({(int64_t)pc.comp_off + off + pay, bs - 8 - pay, total, isize, lo + off});
            total += isize;
            off += bs;
        }
        if (pc.stop < 0) pc.stop = total;          // chunk ends exactly at a block boundary
        pc.n_blk = (int64_t)blocks.size() - pc.first_blk;
        pieces.push_back(pc);
    }
    const double t1 = now();
    if (!b->raw.resize((size_t)total)) return fail(-8, "cannot allocate %lld bytes", (long long)total);
    {
        std::atomic<int> bad(0);
        std::atomic<int64_t> next(0);
        std::vector<std::thread> th;
        const int nt = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, (int64_t)blocks.size() / 4));
        for (int t = 0; t < nt; ++t)
            th.emplace_back([&] {
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
                for (;;) {
                    const int64_t k0 = next.fetch_add(8);
                    if (k0 >= (int64_t)blocks.size()) break;
                    const int64_t k1 = std::min<int64_t>(blocks.size(), k0 + 8);
                    for (int64_t k = k0; k < k1; ++k)
                        if (inflate_one(zs, comp.data() + blocks[k].in_off, blocks[k].in_len, b->raw.data() + blocks[k].out_off, blocks[k].out_len) != 0) bad = 1;
                }
                inflateEnd(&zs);
            });
        for (auto &x : th) x.join();
        if (bad) return fail(-3, "corrupt BGZF block in %s", path);
    }
    const double t2 = now();
    // ---- records of each chunk that overlap the region (pos < end && bam_endpos > beg)
    const uint8_t *r = b->raw.data();
    int maxname = 1;
    for (const Piece &pc : pieces) {
        int64_t p = pc.start;
        while (p < pc.stop) {
            if (p + 36 > total) return fail(-6, "index %s does not match %s (record beyond the chunk)", bai.c_str(), path);
            const int32_t bs = rdi32(r + p);
            if (bs < 32 || p + 4 + bs > total) return fail(-6, "index %s does not match %s (bad record size)", bai.c_str(), path);
            const uint8_t *q = r + p + 4;
            const int l_name = q[8], n_cig = rd16(q + 12);
            const int32_t l_seq = rdi32(q + 16), rtid = rdi32(q), pos = rdi32(q + 4);
            if (l_seq < 0 || 32 + (int64_t)l_name + 4ll * n_cig + (l_seq + 1) / 2 + l_seq > bs)
                return fail(-4, "inconsistent BAM record in %s", path);
            p += 4 + bs;
            if (rtid != tid) { if (rtid > tid || rtid < 0) break; continue; }
            if (pos >= end) break;                       // coordinate sorted: nothing further can overlap
            int64_t reflen = 0;
            const uint8_t *cig = q + 32 + l_name;
            for (int k = 0; k < n_cig; ++k) {
                const uint32_t w = rd32(cig + 4 * k);
                if (REF_CONSUMING[w & 15]) reflen += (w >> 4);
            }
            if (rd16(q + 14) & 4) reflen = 0;
            if ((int64_t)pos + (reflen > 0 ? reflen : 1) <= beg) continue;
            b->rec.push_back((int64_t)(q - r));
            b->cig_off.push_back(b->cig_off.back() + n_cig);
            b->seq_off.push_back(b->seq_off.back() + (l_seq + 1) / 2);
            b->qual_off.push_back(b->qual_off.back() + l_seq);
            maxname = std::max(maxname, l_name > 0 ? l_name - 1 : 0);
        }
    }
    b->name_width = std::max(1, maxname);
    b->t_read = t1 - t0; b->t_inflate = t2 - t1; b->t_index = now() - t2;
    guard.b = nullptr;
    *out = b;
    return 0;
}

// Builds the BAI index of a coordinate-sorted BAM (what `samtools index` writes: binning index with
// merged chunks per bin, 16 kb linear index, the per-reference pseudo-bin 37450 and n_no_coor).
extern "C" int pss_bam_index_build(const char *path, const char *bai_path, int n_threads) {
    pss_bam *b = nullptr;
    int rc = pss_bam_open(path, n_threads, &b);
    if (rc != 0) return rc;
    struct Guard { pss_bam *b; ~Guard() { delete b; } } guard{b};
    const uint8_t *r = b->raw.data();
    const int64_t n = (int64_t)b->rec.size();
    const int32_t n_ref = (int32_t)b->ref_names.size();
    // virtual offset of a position of the inflated stream
    size_t bk = 0;
    auto voff = [&](int64_t p) -> uint64_t {      // p non-decreasing across calls
        while (bk + 1 < b->blocks.size() && (p >= b->blocks[bk].out_off + b->blocks[bk].out_len)) ++bk;
        return ((uint64_t)b->blocks[bk].file_off << 16) | (uint64_t)(p - b->blocks[bk].out_off);
    };
    struct RefIdx {
        std::vector<std::pair<uint32_t, Chunk>> chunks;    // (bin, chunk) in file order
        std::vector<uint64_t> lin;
        uint64_t off_beg = 0, off_end = 0, n_mapped = 0, n_unmapped = 0;
        bool any = false;
    };
    std::vector<RefIdx> idx((size_t)n_ref);
    uint64_t n_no_coor = 0;
    int32_t last_tid = -1, last_pos = -1;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *q = r + b->rec[i];
        const int32_t tid = rdi32(q), pos = rdi32(q + 4), bs = rdi32(q - 4);
        const uint16_t flag = rd16(q + 14);
        const uint64_t v0 = voff(b->rec[i] - 4), v1 = voff(b->rec[i] + bs);
        if (tid < 0) { ++n_no_coor; continue; }
        if (tid >= n_ref) return fail(-4, "record with reference %d beyond the header", tid);
        if (tid < last_tid || (tid == last_tid && pos < last_pos)) return fail(-7, "BAM file is not coordinate sorted: %s", path);
        last_tid = tid; last_pos = pos;
        const int l_name = q[8], n_cig = rd16(q + 12);
        int64_t reflen = 0;
        const uint8_t *cig = q + 32 + l_name;
        for (int k = 0; k < n_cig; ++k) {
            const uint32_t w = rd32(cig + 4 * k);
            if (REF_CONSUMING[w & 15]) reflen += (w >> 4);
        }
        if (flag & 4) reflen = 0;
        const int64_t beg = pos < 0 ? 0 : pos, end = beg + (reflen > 0 ? reflen : 1);
        RefIdx &R = idx[(size_t)tid];
        if (!R.any) { R.off_beg = v0; R.any = true; }
        R.off_end = v1;
        if (flag & 4) ++R.n_unmapped; else ++R.n_mapped;
        const uint32_t bin = (uint32_t)reg2bin(beg, end);
        if (!R.chunks.empty() && R.chunks.back().first == bin && R.chunks.back().second.end == v0) R.chunks.back().second.end = v1;
        else R.chunks.push_back({bin, {v0, v1}});
        const int64_t w0 = beg >> 14, w1 = (end - 1) >> 14;
        if ((int64_t)R.lin.size() <= w1) R.lin.resize((size_t)w1 + 1, 0);
        for (int64_t w = w0; w <= w1; ++w) if (R.lin[(size_t)w] == 0) R.lin[(size_t)w] = v0;
    }
    std::vector<uint8_t> o;
    o.insert(o.end(), {'B', 'A', 'I', 1});
    auto w64 = [&](uint64_t x) { for (int k = 0; k < 8; ++k) o.push_back((uint8_t)(x >> (8 * k))); };
    wr32(o, (uint32_t)n_ref);
    for (RefIdx &R : idx) {
        std::stable_sort(R.chunks.begin(), R.chunks.end(), [](const std::pair<uint32_t, Chunk> &a, const std::pair<uint32_t, Chunk> &c) { return a.first < c.first; });
        int32_t n_bin = 0;
        for (size_t k = 0; k < R.chunks.size(); ++k) if (k == 0 || R.chunks[k].first != R.chunks[k - 1].first) ++n_bin;
        wr32(o, (uint32_t)(n_bin + (R.any ? 1 : 0)));
        for (size_t k = 0; k < R.chunks.size();) {
            size_t e = k;
            while (e < R.chunks.size() && R.chunks[e].first == R.chunks[k].first) ++e;
            wr32(o, R.chunks[k].first);
            wr32(o, (uint32_t)(e - k));
            for (size_t c = k; c < e; ++c) { w64(R.chunks[c].second.beg); w64(R.chunks[c].second.end); }
            k = e;
        }
        if (R.any) {
            wr32(o, 37450u); wr32(o, 2u);
            w64(R.off_beg); w64(R.off_end); w64(R.n_mapped); w64(R.n_unmapped);
        }
        // empty windows inherit the previous entry, as htslib writes them
        for (size_t w = 1; w < R.lin.size(); ++w) if (R.lin[w] == 0) R.lin[w] = R.lin[w - 1];
        wr32(o, (uint32_t)R.lin.size());
        for (uint64_t v : R.lin) w64(v);
    }
    w64(n_no_coor);
    std::string bai = bai_path ? bai_path : std::string(path) + ".bai";
    FILE *fh = fopen(bai.c_str(), "wb");
    if (!fh) return fail(-2, "cannot create %s", bai.c_str());
    fwrite(o.data(), 1, o.size(), fh);
    if (fclose(fh) != 0) return fail(-2, "write error on %s", bai.c_str());
    return 0;
}

extern "C" void pss_bam_close(pss_bam *b) { delete b; }
extern "C" int32_t pss_bam_n_ref(const pss_bam *b) { return b ? (int32_t)b->ref_names.size() : 0; }
extern "C" const char *pss_bam_ref_name(const pss_bam *b, int32_t i) {
    return (b && i >= 0 && i < (int32_t)b->ref_names.size()) ? b->ref_names[i].c_str() : "";
}
extern "C" int64_t pss_bam_ref_len(const pss_bam *b, int32_t i) {
    return (b && i >= 0 && i < (int32_t)b->ref_lens.size()) ? b->ref_lens[i] : -1;
}

extern "C" int pss_bam_sizes(const pss_bam *b, int64_t *n, int64_t *n_cigar, int64_t *n_seq, int64_t *n_qual,
                             int32_t *name_width) {
    if (!b) return fail(-1, "NULL handle");
    if (n) *n = (int64_t)b->rec.size();
    if (n_cigar) *n_cigar = b->cig_off.back();
    if (n_seq) *n_seq = b->seq_off.back();
    if (n_qual) *n_qual = b->qual_off.back();
    if (name_width) *name_width = b->name_width;
    return 0;
}

extern "C" int pss_bam_timing(const pss_bam *b, double *t_read, double *t_inflate, double *t_index, int64_t *raw_bytes) {
    if (!b) return fail(-1, "NULL handle");
    if (t_read) *t_read = b->t_read;
    if (t_inflate) *t_inflate = b->t_inflate;
    if (t_index) *t_index = b->t_index;
    if (raw_bytes) *raw_bytes = (int64_t)b->raw.size();
    return 0;
}

extern "C" int pss_bam_fill(const pss_bam *b, const pss_host_arrays *a) {
    if (!b || !a) return fail(-1, "NULL argument");
    const uint8_t *r = b->raw.data();
    const int64_t n = (int64_t)b->rec.size();
    const int W = b->name_width;
    parallel_for(n, b->n_threads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t i = lo; i < hi; ++i) {
            const uint8_t *q = r + b->rec[i];
            const int l_name = q[8], n_cig = rd16(q + 12);
            const int32_t l_seq = rdi32(q + 16), pos = rdi32(q + 4);
            const uint16_t flag = rd16(q + 14);
            if (a->tid) a->tid[i] = rdi32(q);
            if (a->pos) a->pos[i] = pos;
            if (a->mapq) a->mapq[i] = q[9];
            if (a->flag) a->flag[i] = flag;
            if (a->l_seq) a->l_seq[i] = l_seq;
            if (a->mtid) a->mtid[i] = rdi32(q + 20);
            if (a->mpos) a->mpos[i] = rdi32(q + 24);
            if (a->tlen) a->tlen[i] = rdi32(q + 28);
            const uint8_t *name = q + 32, *cig = name + l_name, *seq = cig + 4 * n_cig, *qual = seq + (l_seq + 1) / 2;
            if (a->names) {
                uint8_t *dst = a->names + (size_t)i * W;
                int ln = l_name > 0 ? l_name - 1 : 0;
                memcpy(dst, name, (size_t)ln);
                memset(dst + ln, 0, (size_t)(W - ln));
            }
            int64_t reflen = 0;
            if (a->cigar) memcpy(a->cigar + b->cig_off[i], cig, (size_t)4 * n_cig);
            for (int k = 0; k < n_cig; ++k) {
                uint32_t w = rd32(cig + 4 * k);
                if (REF_CONSUMING[w & 15]) reflen += (w >> 4);
            }
            if (flag & 4) reflen = 0;
            if (a->endpos) a->endpos[i] = (int64_t)pos + (reflen > 0 ? reflen : 1);
            if (a->seq4) memcpy(a->seq4 + b->seq_off[i], seq, (size_t)(l_seq + 1) / 2);
            if (a->qual) memcpy(a->qual + b->qual_off[i], qual, (size_t)l_seq);
            if (a->cigar_off) a->cigar_off[i] = b->cig_off[i];
            if (a->seq_off) a->seq_off[i] = b->seq_off[i];
            if (a->qual_off) a->qual_off[i] = b->qual_off[i];
        }
    });
    if (a->cigar_off) a->cigar_off[n] = b->cig_off[n];
    if (a->seq_off) a->seq_off[n] = b->seq_off[n];
    if (a->qual_off) a->qual_off[n] = b->qual_off[n];
    return 0;
}

// ------------------------------------------------------------------------------ row packing
//
// load_* returns a packed structured array (one record per genome position, e.g. 249 bytes for
// variation_strand); the engine produces columns.  numpy fills such an array one strided field at
// a time (0.7 M rows/s for 63 fields); this packs blocks of 128 rows -- the block of the output
// stays in L1 while the columns are read sequentially -- on all host threads.

extern "C" int pss_pack_rows(uint8_t *dst, int64_t n_rows, int32_t row_bytes, int32_t n_cols, const void *const *src,
                             const int32_t *src_stride, const int32_t *n_bytes, const int32_t *dst_off, int n_threads) {
    if (n_rows < 0 || row_bytes < 1 || n_cols < 0 || (n_cols > 0 && (!src || !src_stride || !n_bytes || !dst_off)) || (n_rows > 0 && !dst))
        return fail(-1, "bad argument");
    for (int c = 0; c < n_cols; ++c)
        if (!src[c] || n_bytes[c] < 0 || dst_off[c] < 0 || dst_off[c] + n_bytes[c] > row_bytes || src_stride[c] < 0)
            return fail(-1, "bad column %d", c);
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int64_t BLK = 128;
    const int64_t nblk = (n_rows + BLK - 1) / BLK;
    parallel_for(nblk, (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, nblk / 8)), [&](int64_t b0, int64_t b1, int) {
        for (int64_t b = b0; b < b1; ++b) {
            const int64_t r0 = b * BLK, r1 = std::min(n_rows, r0 + BLK);
            uint8_t *out = dst + (size_t)r0 * (size_t)row_bytes;
            for (int c = 0; c < n_cols; ++c) {
                const uint8_t *in = (const uint8_t *)src[c] + (size_t)r0 * (size_t)src_stride[c];
                uint8_t *o = out + dst_off[c];
                const int st = src_stride[c], nb = n_bytes[c];
                if (nb == 4) {
                    for (int64_t r = r0; r < r1; ++r, in += st, o += row_bytes) { uint32_t v; memcpy(&v, in, 4); memcpy(o, &v, 4); }
                } else if (nb == 1) {
                    for (int64_t r = r0; r < r1; ++r, in += st, o += row_bytes) *o = *in;
                } else {
                    for (int64_t r = r0; r < r1; ++r, in += st, o += row_bytes) memcpy(o, in, (size_t)nb);
                }
            }
        }
    });
    return 0;
}

// dst[c][k] = src[c][idx[k]] for n_cols int32 columns (rows of [n_cols][*] arrays, strides in elements):
// the compaction of the engine's per-position columns to the rows htslib emits, on all host threads.
extern "C" int pss_take_cols_i32(const int32_t *src, int64_t src_stride, int32_t n_cols, const int64_t *idx, int64_t n,
                                 int32_t *dst, int64_t dst_stride, int n_threads) {
    if (n < 0 || n_cols < 0 || (n > 0 && n_cols > 0 && (!src || !idx || !dst))) return fail(-1, "bad argument");
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int64_t BLK = 4096;
    const int64_t nblk = (n + BLK - 1) / BLK;
    parallel_for(nblk * n_cols, (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, nblk * n_cols / 4)), [&](int64_t t0, int64_t t1, int) {
        for (int64_t t = t0; t < t1; ++t) {
            const int64_t c = t / nblk, b = t - c * nblk;
            const int64_t k0 = b * BLK, k1 = std::min(n, k0 + BLK);
            const int32_t *in = src + c * src_stride;
            int32_t *out = dst + c * dst_stride;
            for (int64_t k = k0; k < k1; ++k) out[k] = in[idx[k]];
        }
    });
    return 0;
}

// One raw DEFLATE stream, whole: mode 0 = the decoder of the BAM path (own decoder, zlib for what it declines),
// 1 = own decoder only (-1 when it declines), 2 = zlib only.  For tests and timing.
extern "C" int pss_inflate_raw(const uint8_t *src, int64_t in_len, uint8_t *dst, int64_t out_len, int mode) {
    if (!src || !dst || in_len < 0 || out_len < 0) return fail(-1, "bad argument");
    if (mode == 1) return pss_inflate_fast(src, (size_t)in_len, dst, (size_t)out_len);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return fail(-3, "zlib init failed");
    int rc;
    if (mode == 2) {
        zs.next_in = (Bytef *)src; zs.avail_in = (uInt)in_len;
        zs.next_out = dst; zs.avail_out = (uInt)out_len;
        rc = (out_len == 0 || (inflate(&zs, Z_FINISH) == Z_STREAM_END && zs.avail_out == 0)) ? 0 : -1;
    } else rc = inflate_one(zs, src, in_len, dst, out_len);
    inflateEnd(&zs);
    return rc;
}

// ------------------------------------------------------------------------------ text rows
//
// write_csv (pysamstats/io.py:14-88) pushes one dict per row through csv.writer; here a block of rows is
// formatted from the columns on all host threads: decimal integers, a constant (already quoted if it had to
// be) chrom text, one-character ref bases, joined by the dialect's delimiter and line terminator -- the bytes
// csv.writer writes for values that need no quoting (the caller keeps quoting cases to itself).

static inline int dec_digits(uint64_t u) {
    if (u < 100) return u < 10 ? 1 : 2;
    if (u < 10000) return u < 1000 ? 3 : 4;
    if (u < 1000000) return u < 100000 ? 5 : 6;
    if (u < 100000000) return u < 10000000 ? 7 : 8;
    int n = 9;
    for (uint64_t t = 1000000000ull; n < 20 && u >= t; t *= 10) ++n;
    return n;
}

static inline int dec_len(int64_t v) {
    return v < 0 ? 1 + dec_digits((uint64_t)0 - (uint64_t)v) : dec_digits((uint64_t)v);
}

static inline char *dec_put(char *p, int64_t v) {
    static const char lut[201] =
        "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
        "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
        "8081828384858687888990919293949596979899";
    uint64_t u = (uint64_t)v;
    if (v < 0) { *p++ = '-'; u = (uint64_t)0 - u; }
    const int n = dec_digits(u);
    char *e = p + n;
    while (u >= 100) { const uint64_t q = u / 100; e -= 2; e[0] = lut[2 * (u - q * 100)]; e[1] = lut[2 * (u - q * 100) + 1]; u = q; }
    if (u >= 10) { e[-2] = lut[2 * u]; e[-1] = lut[2 * u + 1]; }
    else e[-1] = (char)('0' + u);
    return p + n;
}

// kind[c]: 0 constant text (src = bytes, width = length), 1 int32 column, 2 int64 column, 3 one-byte text
// column (a NUL byte is the empty string); src_stride in bytes.  *out is malloc'ed (pss_host_free).
extern "C" int pss_format_rows(int64_t n_rows, int32_t n_cols, const int32_t *kind, const void *const *src,
                               const int32_t *src_stride, const int32_t *width, const char *delimiter, int32_t delimiter_len,
                               const char *terminator, int32_t terminator_len, int n_threads, char **out, int64_t *out_len) {
    if (n_rows < 0 || n_cols < 1 || !kind || !src || !src_stride || !width || !out || !out_len || delimiter_len < 0 ||
        terminator_len < 0 || (delimiter_len > 0 && !delimiter) || (terminator_len > 0 && !terminator))
        return fail(-1, "bad argument");
    for (int c = 0; c < n_cols; ++c)
        if (kind[c] < 0 || kind[c] > 3 || !src[c] || (kind[c] == 0 && width[c] < 0)) return fail(-1, "bad column %d", c);
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int64_t BLK = 4096;
    const int64_t nblk = (n_rows + BLK - 1) / BLK;
    std::vector<int64_t> off((size_t)nblk + 1, 0);
    auto rows_of = [&](int64_t b, auto &&cell) {
        const int64_t r0 = b * BLK, r1 = std::min(n_rows, r0 + BLK);
        for (int64_t r = r0; r < r1; ++r)
            for (int c = 0; c < n_cols; ++c) cell(r, c);
    };
    const int nt = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, nblk / 2));
    parallel_for(nblk, nt, [&](int64_t b0, int64_t b1, int) {
        for (int64_t b = b0; b < b1; ++b) {
            int64_t len = 0;
            rows_of(b, [&](int64_t r, int c) {
                const uint8_t *p = (const uint8_t *)src[c] + (size_t)r * (size_t)src_stride[c];
                switch (kind[c]) {
                    case 0: len += width[c]; break;
                    case 1: { int32_t v; memcpy(&v, p, 4); len += dec_len(v); break; }
                    case 2: { int64_t v; memcpy(&v, p, 8); len += dec_len(v); break; }
                    default: len += *p != 0;
                }
            });
            const int64_t nr = std::min(n_rows, (b + 1) * BLK) - b * BLK;
            off[(size_t)b + 1] = len + nr * ((int64_t)(n_cols - 1) * delimiter_len + terminator_len);
        }
    });
    for (int64_t b = 0; b < nblk; ++b) off[(size_t)b + 1] += off[(size_t)b];
    const int64_t total = off[(size_t)nblk];
    char *buf = (char *)malloc((size_t)std::max<int64_t>(total, 1));
    if (!buf) return fail(-1, "out of memory (%lld bytes of text)", (long long)total);
    parallel_for(nblk, nt, [&](int64_t b0, int64_t b1, int) {
        for (int64_t b = b0; b < b1; ++b) {
            char *o = buf + off[(size_t)b];
            rows_of(b, [&](int64_t r, int c) {
                const uint8_t *p = (const uint8_t *)src[c] + (size_t)r * (size_t)src_stride[c];
                switch (kind[c]) {
                    case 0: memcpy(o, src[c], (size_t)width[c]); o += width[c]; break;
                    case 1: { int32_t v; memcpy(&v, p, 4); o = dec_put(o, v); break; }
                    case 2: { int64_t v; memcpy(&v, p, 8); o = dec_put(o, v); break; }
                    default: if (*p) *o++ = (char)*p;
                }
                if (c + 1 < n_cols) {
                    if (delimiter_len == 1) *o++ = delimiter[0];
                    else { memcpy(o, delimiter, (size_t)delimiter_len); o += delimiter_len; }
                } else {
                    for (int k = 0; k < terminator_len; ++k) *o++ = terminator[k];
                }
            });
        }
    });
    *out = buf;
    *out_len = total;
    return 0;
}

extern "C" void pss_host_free(void *p) { free(p); }

// modes 3 / 4 of the probe below: the same stores as whole 64-byte lines (AVX-512)
__attribute__((target("avx512f,avx512bw"))) static void bw_avx512(uint8_t *b, const uint8_t *sp, int64_t bytes, int reps) {
    const int64_t n4 = bytes / 4;
    for (int r = 0; r < reps; ++r) {
        int *d = (int *)b;
        if (!sp) {
            const __m512i v = _mm512_set1_epi32(r);
            for (int64_t i = 0; i < n4; i += 16) _mm512_stream_si512((__m512i *)(d + i), v);
        } else {
            for (int64_t i = 0; i < n4; i += 16)
                _mm512_stream_si512((__m512i *)(d + i), _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(sp + i))));
        }
    }
    _mm_sfence();
}

// Host memory bandwidth probe: `n_threads` threads stream over a private buffer of `bytes` each, `reps` times.
// mode 0: non-temporal int32 stores (what the result path does to the caller's columns), 1: widen u8 -> int32
// (1 byte read, 4 written), 2: plain reads.  Returns GB/s of bytes WRITTEN (mode 2: read), < 0 on failure.
extern "C" double pss_host_bandwidth(int64_t bytes, int n_threads, int reps, int mode) {
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    if (bytes < 4096 || reps < 1) return -1.0;
    bytes &= ~(int64_t)63;
    std::vector<uint8_t *> bufs((size_t)n_threads, nullptr), srcs((size_t)n_threads, nullptr);
    for (int t = 0; t < n_threads; ++t) {
        if (posix_memalign((void **)&bufs[(size_t)t], 64, (size_t)bytes) != 0) return -1.0;
        if ((mode == 1 || mode == 4) && posix_memalign((void **)&srcs[(size_t)t], 64, (size_t)bytes / 4) != 0) return -1.0;
    }
    std::atomic<int> ready(0), go(0);
    std::atomic<uint64_t> sink(0);
    std::vector<std::thread> th;
    if ((mode == 3 || mode == 4) && !__builtin_cpu_supports("avx512f")) return -2.0;
    auto body = [&](int t, int r_n) {
        if (mode == 3 || mode == 4) { bw_avx512(bufs[(size_t)t], mode == 4 ? srcs[(size_t)t] : nullptr, bytes, r_n); return; }
        uint8_t *b = bufs[(size_t)t];
        const int64_t n4 = bytes / 4;
        for (int r = 0; r < r_n; ++r) {
            if (mode == 0) {
                int *d = (int *)b;
                const __m128i v = _mm_set1_epi32(r);
                for (int64_t i = 0; i < n4; i += 4) _mm_stream_si128((__m128i *)(d + i), v);
            } else if (mode == 1) {
                const uint8_t *sp = srcs[(size_t)t];
                int *d = (int *)b;
                for (int64_t i = 0; i < n4; i += 4) {
                    uint32_t w4;
                    memcpy(&w4, sp + i, 4);
                    const __m128i w = _mm_cvtsi32_si128((int)w4), z = _mm_setzero_si128();
                    _mm_stream_si128((__m128i *)(d + i), _mm_unpacklo_epi16(_mm_unpacklo_epi8(w, z), z));
                }
            } else {
                const uint64_t *q = (const uint64_t *)b;
                uint64_t acc = 0;
                for (int64_t i = 0; i < bytes / 8; ++i) acc += q[i];
                sink += acc;
            }
        }
    };
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            memset(bufs[(size_t)t], 1, (size_t)bytes);                 // first touch outside the timed part
            if (mode == 1 || mode == 4) memset(srcs[(size_t)t], 3, (size_t)bytes / 4);
            ready.fetch_add(1);
            while (!go.load()) std::this_thread::yield();
            body(t, reps);
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    const double t0 = now();
    go.store(1);
    for (auto &x : th) x.join();
    const double dt = now() - t0;
    for (auto p2 : bufs) free(p2);
    for (auto p2 : srcs) free(p2);
    return dt > 0 ? (double)bytes * n_threads * reps / dt / 1e9 : -1.0;
}

// ------------------------------------------------------------------------------ name_prev

extern "C" int pss_name_prev(const uint8_t *names, int64_t n, int32_t width, int32_t *prev) {
    if (n < 0 || width < 1 || (n > 0 && (!names || !prev))) return fail(-1, "bad argument");
    if (n > 0x7fffffff) return fail(-1, "too many records");
    size_t cap = 16;
    while (cap < (size_t)n * 2) cap <<= 1;
    std::vector<int32_t> table(cap, -1);          // last row index seen per slot
    const size_t mask = cap - 1;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *s = names + (size_t)i * width;
        uint64_t h = 1469598103934665603ull;      // FNV-1a
        for (int k = 0; k < width && s[k]; ++k) { h ^= s[k]; h *= 1099511628211ull; }
        size_t slot = (size_t)(h ^ (h >> 29)) & mask;
        int32_t found = -1;
        for (;;) {
            int32_t j = table[slot];
            if (j < 0) break;
            if (memcmp(names + (size_t)j * width, s, (size_t)width) == 0) { found = j; break; }
            slot = (slot + 1) & mask;
        }
        prev[i] = found;
        table[slot] = (int32_t)i;                 // the slot now remembers the LATEST row of this name
    }
    return 0;
}

// ---------------------------------------------------------------------------------- writer

static bool bgzf_block(const uint8_t *src, size_t len, int level, std::vector<uint8_t> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(18 + compressBound((uLong)len) + 8);
    zs.next_in = (Bytef *)src; zs.avail_in = (uInt)len;
    zs.next_out = out.data() + 18; zs.avail_out = (uInt)(out.size() - 26);
    int rc = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    size_t clen = zs.total_out, bsize = 18 + clen + 8;
    if (bsize > 65536) return false;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0};
    memcpy(out.data(), hdr, 16);
    out[16] = (uint8_t)((bsize - 1) & 255); out[17] = (uint8_t)((bsize - 1) >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)len);
    uint8_t *t = out.data() + 18 + clen;
    for (int k = 0; k < 4; ++k) t[k] = (crc >> (8 * k)) & 255;
    for (int k = 0; k < 4; ++k) t[4 + k] = ((uint32_t)len >> (8 * k)) & 255;
    out.resize(bsize);
    return true;
}

extern "C" int pss_bam_write(const char *path, int32_t n_ref, const char *const *ref_names, const int64_t *ref_lens,
                             int64_t n, const pss_host_arrays *a, int32_t name_width, int level, int n_threads) {
    if (!path || !a || n < 0 || n_ref < 0) return fail(-1, "bad argument");
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    // uncompressed stream, built in record ranges by the threads
    std::vector<uint8_t> head;
    head.insert(head.end(), {'B', 'A', 'M', 1});
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (int i = 0; i < n_ref; ++i) text += std::string("@SQ\tSN:") + ref_names[i] + "\tLN:" + std::to_string(ref_lens[i]) + "\n";
    wr32(head, (uint32_t)text.size());
    head.insert(head.end(), text.begin(), text.end());
    wr32(head, (uint32_t)n_ref);
    for (int i = 0; i < n_ref; ++i) {
        size_t ln = strlen(ref_names[i]) + 1;
        wr32(head, (uint32_t)ln);
        head.insert(head.end(), ref_names[i], ref_names[i] + ln);
        wr32(head, (uint32_t)ref_lens[i]);
    }
    std::vector<int64_t> roff((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *nm = a->names + (size_t)i * name_width;
        int ln = 0;
        while (ln < name_width && nm[ln]) ++ln;
        int64_t ncig = a->cigar_off[i + 1] - a->cigar_off[i];
        int64_t sz = 36 + (ln + 1) + 4 * ncig + (a->seq_off[i + 1] - a->seq_off[i]) + a->l_seq[i];
        roff[i + 1] = roff[i] + sz;
    }
    std::vector<uint8_t> raw(head.size() + (size_t)roff[n]);
    memcpy(raw.data(), head.data(), head.size());
    uint8_t *body = raw.data() + head.size();
    parallel_for(n, n_threads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t i = lo; i < hi; ++i) {
            uint8_t *q = body + roff[i];
            const uint8_t *nm = a->names + (size_t)i * name_width;
            int ln = 0;
            while (ln < name_width && nm[ln]) ++ln;
            const int64_t ncig = a->cigar_off[i + 1] - a->cigar_off[i];
            const int32_t l_seq = a->l_seq[i];
            const int64_t nsb = a->seq_off[i + 1] - a->seq_off[i];
            int32_t bs = (int32_t)(roff[i + 1] - roff[i] - 4);
            int64_t end = a->endpos ? a->endpos[i] : (int64_t)a->pos[i] + 1;
            int32_t v[8] = {bs, a->tid[i], a->pos[i], 0, 0, l_seq, a->mtid[i], a->mpos ? a->mpos[i] : -1};
            memcpy(q, &v[0], 12);
            q[12] = (uint8_t)(ln + 1); q[13] = a->mapq[i];
            uint16_t bin = (uint16_t)reg2bin(a->pos[i], end);
            memcpy(q + 14, &bin, 2);
            uint16_t nc16 = (uint16_t)ncig, fl = a->flag[i];
            memcpy(q + 16, &nc16, 2); memcpy(q + 18, &fl, 2);
            memcpy(q + 20, &v[5], 4); memcpy(q + 24, &v[6], 4); memcpy(q + 28, &v[7], 4);
            int32_t tl = a->tlen[i];
            memcpy(q + 32, &tl, 4);
            uint8_t *w = q + 36;
            memcpy(w, nm, (size_t)ln); w[ln] = 0; w += ln + 1;
            memcpy(w, a->cigar + a->cigar_off[i], (size_t)4 * ncig); w += 4 * ncig;
            memcpy(w, a->seq4 + a->seq_off[i], (size_t)nsb); w += nsb;
            if (a->qual && a->qual_off) memcpy(w, a->qual + a->qual_off[i], (size_t)l_seq);
            else memset(w, 0xff, (size_t)l_seq);
        }
    });
    // BGZF blocks of <= 0xf000 payload bytes, compressed in parallel, written in order
    const size_t BLK = 0xf000;   // incompressible payloads must still fit a 64 KiB block
    const int64_t nblk = (int64_t)((raw.size() + BLK - 1) / BLK);
    std::vector<std::vector<uint8_t>> outb((size_t)nblk);
    std::atomic<int> bad(0);
    parallel_for(nblk, n_threads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t k = lo; k < hi; ++k) {
            size_t o = (size_t)k * BLK, len = std::min(BLK, raw.size() - o);
            if (!bgzf_block(raw.data() + o, len, level, outb[(size_t)k])) bad = 1;
        }
    });
    if (bad) return fail(-5, "deflate failed");
    FILE *fh = fopen(path, "wb");
    if (!fh) return fail(-2, "cannot create %s", path);
    for (auto &blk : outb) fwrite(blk.data(), 1, blk.size(), fh);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, fh);
    if (fclose(fh) != 0) return fail(-2, "write error on %s", path);
    return 0;
}
