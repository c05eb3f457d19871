// Host-side BAM decoder / writer behind include/pss_host.h: multi-threaded BGZF inflate,
// BAM record indexing and parallel structure-of-arrays fill.  Plain C++17 + zlib + pthreads.
//
// BGZF (SAM spec section 4.1): a series of gzip members, each with a "BC" extra subfield
// holding the block size, each inflating to <= 64 KiB, independently decodable -- so blocks
// are distributed over threads.  BAM (section 4.2): after the header, records
// {block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID,
//  next_pos, tlen, read_name, cigar, seq, qual, aux}.
#include "../../include/pss_host.h"

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

struct pss_bam {
    std::vector<uint8_t> raw;            // inflated BAM stream
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
    FILE *fh = fopen(path, "rb");
    if (!fh) return fail(-2, "cannot open %s", path);
    fseek(fh, 0, SEEK_END);
    long long fsize = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    std::vector<uint8_t> comp((size_t)fsize);
    if (fsize > 0 && fread(comp.data(), 1, (size_t)fsize, fh) != (size_t)fsize) { fclose(fh); return fail(-2, "short read on %s", path); }
    fclose(fh);
    double t1 = now();
    // block table
    struct Blk { int64_t in_off, in_len, out_off, out_len; };
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
        blocks.push_back({xend, off + bsize + 1 - 8 - xend, total, isize});
        total += isize;
        off += bsize + 1;
    }
    pss_bam *b = new pss_bam();
    b->n_threads = n_threads;
    b->raw.resize((size_t)total);
    std::atomic<int> bad(0);
    std::atomic<int64_t> next(0);
    {
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
                        if (bl.out_len == 0) continue;
                        inflateReset(&zs);
                        zs.next_in = comp.data() + bl.in_off; zs.avail_in = (uInt)bl.in_len;
                        zs.next_out = b->raw.data() + bl.out_off; zs.avail_out = (uInt)bl.out_len;
                        int rc = inflate(&zs, Z_FINISH);
                        if (rc != Z_STREAM_END || zs.avail_out != 0) bad = 1;
                    }
                }
                inflateEnd(&zs);
            });
        for (auto &x : th) x.join();
    }
    if (bad) { delete b; return fail(-3, "corrupt BGZF block in %s", path); }
    std::vector<uint8_t>().swap(comp);
    double t2 = now();
    const uint8_t *r = b->raw.data();
    const int64_t n = (int64_t)b->raw.size();
    if (n < 12 || memcmp(r, "BAM\1", 4) != 0) { delete b; return fail(-4, "not a BAM file: %s", path); }
    int64_t p = 8 + rdi32(r + 4);
    if (p + 4 > n) { delete b; return fail(-4, "truncated BAM header"); }
    int32_t n_ref = rdi32(r + p);
    p += 4;
    for (int32_t i = 0; i < n_ref; ++i) {
        if (p + 4 > n) { delete b; return fail(-4, "truncated BAM header"); }
        int32_t l_name = rdi32(r + p);
        if (l_name < 1 || p + 8 + l_name > n) { delete b; return fail(-4, "truncated BAM header"); }
        b->ref_names.emplace_back((const char *)(r + p + 4), (size_t)(l_name - 1));
        b->ref_lens.push_back(rdi32(r + p + 4 + l_name));
        p += 8 + l_name;
    }
    // record index + prefix sums (sequential: one pass over the block_size chain)
    int maxname = 1;
    b->cig_off.push_back(0); b->seq_off.push_back(0); b->qual_off.push_back(0);
    while (p < n) {
        if (p + 36 > n) { delete b; return fail(-4, "truncated BAM record at %lld", (long long)p); }
        int32_t bs = rdi32(r + p);
        if (bs < 32 || p + 4 + bs > n) { delete b; return fail(-4, "bad BAM record size at %lld", (long long)p); }
        const uint8_t *q = r + p + 4;
        int l_name = q[8], n_cig = rd16(q + 12);
        int32_t l_seq = rdi32(q + 16);
        if (l_seq < 0 || 32 + (int64_t)l_name + 4ll * n_cig + (l_seq + 1) / 2 + l_seq > bs) {
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
    b->name_width = std::max(1, maxname);
    b->t_read = t1 - t0; b->t_inflate = t2 - t1; b->t_index = now() - t2;
    *out = b;
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
