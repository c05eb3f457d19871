#include "pss_inflate.h"

#include <string.h>

namespace {

constexpr int LL_BITS = 11, OF_BITS = 8;
constexpr int LL_SIZE = (1 << LL_BITS) + 1024, OF_SIZE = (1 << OF_BITS) + 512;   // primary + room for every subtable
// entry: [31:16] literal / length base / offset base / subtable start, [15:8] extra bits (or subtable index bits),
// [7] literal, [6] subtable pointer, [5] end of block, [3:0] codeword length (0: hole)
constexpr uint32_t F_LIT = 0x80, F_SUB = 0x40, F_EOB = 0x20;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kOffBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                               4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kOffExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t rev_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1); code >>= 1; }
    return r;
}

// kind 0: litlen symbols, 1: offset symbols, 2: code-length symbols (value = symbol)
inline uint32_t make_entry(int kind, int sym, int len) {
    if (kind == 0) {
        if (sym < 256) return ((uint32_t)sym << 16) | F_LIT | (uint32_t)len;
        if (sym == 256) return F_EOB | (uint32_t)len;
        if (sym > 285) return 0;                               // 286, 287 never occur in valid data: a hole
        return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8) | (uint32_t)len;
    }
    if (kind == 1) {
        if (sym > 29) return 0;
        return ((uint32_t)kOffBase[sym] << 16) | ((uint32_t)kOffExtra[sym] << 8) | (uint32_t)len;
    }
    return ((uint32_t)sym << 16) | (uint32_t)len;
}

// canonical Huffman decode table from code lengths; false: over-subscribed (or does not fit the table)
bool build_table(const uint8_t *lens, int n, int kind, int tbits, uint32_t *table, int table_size) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    count[0] = 0;
    uint32_t next[16], code = 0;
    int64_t left = 1;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
        next[l] = code;
    }
    // an incomplete set is an error for zlib (inflate_table) unless it is a single one-bit code or empty: the same here,
    // so that this decoder accepts nothing zlib would reject
    if (left > 0) {
        int maxlen = 0;
        for (int l = 1; l <= 15; ++l) if (count[l]) maxlen = l;
        if (maxlen > 1 || (maxlen == 1 && kind == 2)) return false;
    }
    const int primary = 1 << tbits;
    memset(table, 0, sizeof(uint32_t) * (size_t)primary);
    // longest codeword behind every primary prefix that needs a subtable
    uint8_t sub_max[1 << LL_BITS];
    bool any_long = false;
    for (int l = tbits + 1; l <= 15; ++l) any_long |= count[l] != 0;
    if (any_long) {
        memset(sub_max, 0, (size_t)primary);
        uint32_t nx[16];
        memcpy(nx, next, sizeof(nx));
        for (int i = 0; i < n; ++i) {
            const int l = lens[i];
            if (l <= tbits) { if (l) ++nx[l]; continue; }
            const uint32_t r = rev_bits(nx[l]++, l);
            uint8_t &m = sub_max[r & (uint32_t)(primary - 1)];
            if (l > m) m = (uint8_t)l;
        }
    }
    int used = primary;
    for (int i = 0; i < n; ++i) {
        const int l = lens[i];
        if (!l) continue;
        const uint32_t r = rev_bits(next[l]++, l);
        const uint32_t e = make_entry(kind, i, l);
        if (l <= tbits) {
            for (uint32_t k = r; k < (uint32_t)primary; k += 1u << l) table[k] = e;
            continue;
        }
        const uint32_t pre = r & (uint32_t)(primary - 1);
        const int sbits = sub_max[pre] - tbits;
        if (!(table[pre] & F_SUB)) {
            if (used + (1 << sbits) > table_size) return false;
            memset(table + used, 0, sizeof(uint32_t) << sbits);
            table[pre] = ((uint32_t)used << 16) | ((uint32_t)sbits << 8) | F_SUB | (uint32_t)tbits;
            used += 1 << sbits;
        }
        uint32_t *sub = table + (table[pre] >> 16);
        for (uint32_t k = r >> tbits; k < (1u << sbits); k += 1u << (l - tbits)) sub[k] = e;
    }
    return true;
}

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct FixedTables {
    uint32_t ll[LL_SIZE], of[OF_SIZE];
    FixedTables() {
        uint8_t l[288];
        for (int i = 0; i < 288; ++i) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        build_table(l, 288, 0, LL_BITS, ll, LL_SIZE);
        uint8_t d[32];
        for (int i = 0; i < 32; ++i) d[i] = 5;
        build_table(d, 32, 1, OF_BITS, of, OF_SIZE);
    }
};

}  // namespace

int pss_inflate_fast(const uint8_t *src, size_t in_len, uint8_t *dst, size_t out_len) {
    static const FixedTables fixed;
    const uint8_t *in = src, *const in_end = src + in_len;
    uint8_t *out = dst, *const out_end = dst + out_len;
    uint64_t bitbuf = 0;
    unsigned bitcnt = 0;
    uint32_t ll_dyn[LL_SIZE], of_dyn[OF_SIZE];

    // at least 56 valid bits while 8 input bytes remain; else whatever is left (checked by the consumers)
#define REFILL()                                                                          \
    do {                                                                                  \
        if (in_end - in >= 8) {                                                           \
            bitbuf |= load64(in) << bitcnt;                                               \
            in += (63 - bitcnt) >> 3;                                                     \
            bitcnt |= 56;                                                                 \
        } else                                                                            \
            while (bitcnt <= 56 && in < in_end) { bitbuf |= (uint64_t)*in++ << bitcnt; bitcnt += 8; } \
    } while (0)
#define NEED(n) do { if (bitcnt < (unsigned)(n)) return -1; } while (0)
#define DROP(n) do { bitbuf >>= (n); bitcnt -= (n); } while (0)

    for (;;) {
        REFILL();
        NEED(3);
        const unsigned final_block = (unsigned)bitbuf & 1, type = ((unsigned)bitbuf >> 1) & 3;
        DROP(3);
        const uint32_t *ll, *of;
        if (type == 0) {                                  // stored
            DROP(bitcnt & 7);
            in -= bitcnt >> 3;                            // whole bytes still in the bit buffer go back
            bitbuf = 0; bitcnt = 0;
            if (in_end - in < 4) return -1;
            const unsigned len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
            in += 4;
            if ((len ^ nlen) != 0xffff || (size_t)(in_end - in) < len || (size_t)(out_end - out) < len) return -1;
            memcpy(out, in, len);
            in += len; out += len;
            if (final_block) break;
            continue;
        }
        if (type == 3) return -1;
        if (type == 1) { ll = fixed.ll; of = fixed.of; }
        else {
            NEED(14);
            const unsigned hlit = ((unsigned)bitbuf & 31) + 257, hdist = (((unsigned)bitbuf >> 5) & 31) + 1,
                           hclen = (((unsigned)bitbuf >> 10) & 15) + 4;
            DROP(14);
            if (hlit > 286 || hdist > 30) return -1;
            uint8_t cl[19] = {0};
            for (unsigned i = 0; i < hclen; ++i) {
                if (bitcnt < 3) { REFILL(); NEED(3); }
                cl[kClOrder[i]] = (uint8_t)(bitbuf & 7);
                DROP(3);
            }
            uint32_t cl_table[128];
            if (!build_table(cl, 19, 2, 7, cl_table, 128)) return -1;
            uint8_t lens[286 + 30 + 138];
            unsigned i = 0;
            while (i < hlit + hdist) {
                REFILL();
                const uint32_t e = cl_table[bitbuf & 127];
                const unsigned l = e & 15;
                if (!l) return -1;
                NEED(l);
                DROP(l);
                const unsigned sym = e >> 16;
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                unsigned rep;
                uint8_t v = 0;
                if (sym == 16) { if (!i) return -1; NEED(2); v = lens[i - 1]; rep = 3 + ((unsigned)bitbuf & 3); DROP(2); }
                else if (sym == 17) { NEED(3); rep = 3 + ((unsigned)bitbuf & 7); DROP(3); }
                else { NEED(7); rep = 11 + ((unsigned)bitbuf & 127); DROP(7); }
                if (i + rep > hlit + hdist) return -1;
                memset(lens + i, v, rep);
                i += rep;
            }
            if (!lens[256]) return -1;                    // no end-of-block code
            if (!build_table(lens, (int)hlit, 0, LL_BITS, ll_dyn, LL_SIZE)) return -1;
            if (!build_table(lens + hlit, (int)hdist, 1, OF_BITS, of_dyn, OF_SIZE)) return -1;
            ll = ll_dyn; of = of_dyn;
        }
        // ---- symbols of the block
        for (;;) {
            // fast path: far enough from both ends that no single symbol can run out of input or output, so the
            // only checks left are holes and offsets before the block
            while (in_end - in >= 32 && out_end - out >= 280) {     // three 8-byte loads at in + <= 14; 3 literals + 258 + 15 bytes out
#define REFILL_FAST() do { bitbuf |= load64(in) << bitcnt; in += (63 - bitcnt) >> 3; bitcnt |= 56; } while (0)
                REFILL_FAST();
                uint32_t e = ll[bitbuf & ((1u << LL_BITS) - 1)];
#define PUT_LITS(e) do { *out++ = (uint8_t)((e) >> 16); } while (0)
                if (e & F_LIT) {                          // up to three literals per refill (3 x 15 bits)
                    DROP(e & 15); PUT_LITS(e);
                    e = ll[bitbuf & ((1u << LL_BITS) - 1)];
                    if (e & F_LIT) {
                        DROP(e & 15); PUT_LITS(e);
                        e = ll[bitbuf & ((1u << LL_BITS) - 1)];
                        if (e & F_LIT) { DROP(e & 15); PUT_LITS(e); continue; }
                    }
                    REFILL_FAST();
                    e = ll[bitbuf & ((1u << LL_BITS) - 1)];
                }
                if (e & F_SUB) e = ll[(e >> 16) + ((bitbuf >> LL_BITS) & ((1u << ((e >> 8) & 0xff)) - 1))];
                unsigned l = e & 15;
                if (!l) return -1;
                DROP(l);
                if (e & F_LIT) { PUT_LITS(e); continue; }
                if (e & F_EOB) goto block_done;
                unsigned x = (e >> 8) & 0xff;
                const unsigned length = (e >> 16) + ((unsigned)bitbuf & ((1u << x) - 1));
                DROP(x);
                REFILL_FAST();
                e = of[bitbuf & ((1u << OF_BITS) - 1)];
                if (e & F_SUB) e = of[(e >> 16) + ((bitbuf >> OF_BITS) & ((1u << ((e >> 8) & 0xff)) - 1))];
                l = e & 15;
                if (!l) return -1;
                DROP(l);
                x = (e >> 8) & 0xff;
                const size_t dist = (e >> 16) + ((unsigned)bitbuf & ((1u << x) - 1));
                DROP(x);
                if (dist > (size_t)(out - dst)) return -1;
                const uint8_t *from = out - dist;
                uint8_t *o = out;
                out += length;
                if (dist >= 16) {
                    do { memcpy(o, from, 16); o += 16; from += 16; } while (o < out);
                } else if (dist >= 8) {
                    do { memcpy(o, from, 8); o += 8; from += 8; } while (o < out);
                } else if (dist == 1) {
                    memset(o, *from, length);
                } else {
                    do { *o++ = *from++; } while (o < out);
                }
#undef REFILL_FAST
#undef PUT_LITS
            }
            REFILL();
            uint32_t e = ll[bitbuf & ((1u << LL_BITS) - 1)];
            if (e & F_SUB) e = ll[(e >> 16) + ((bitbuf >> LL_BITS) & ((1u << ((e >> 8) & 0xff)) - 1))];
            unsigned l = e & 15;
            if (!l || l > bitcnt) return -1;
            DROP(l);
            if (e & F_LIT) {
                if (out == out_end) return -1;
                *out++ = (uint8_t)(e >> 16);
                continue;
            }
            if (e & F_EOB) break;
            unsigned x = (e >> 8) & 0xff;
            if (x > bitcnt) return -1;
            const unsigned length = (e >> 16) + ((unsigned)bitbuf & ((1u << x) - 1));
            DROP(x);
            if (bitcnt < 28) REFILL();                    // offset code (<= 15) + extra bits (<= 13)
            e = of[bitbuf & ((1u << OF_BITS) - 1)];
            if (e & F_SUB) e = of[(e >> 16) + ((bitbuf >> OF_BITS) & ((1u << ((e >> 8) & 0xff)) - 1))];
            l = e & 15;
            if (!l || l > bitcnt) return -1;
            DROP(l);
            x = (e >> 8) & 0xff;
            if (x > bitcnt) return -1;
            const size_t dist = (e >> 16) + ((unsigned)bitbuf & ((1u << x) - 1));
            DROP(x);
            if (dist > (size_t)(out - dst) || length > (size_t)(out_end - out)) return -1;
            const uint8_t *from = out - dist;
            if (dist >= 8 && (size_t)(out_end - out) >= length + 8) {
                uint8_t *o = out;
                const uint8_t *const stop = out + length;
                do { memcpy(o, from, 8); o += 8; from += 8; } while (o < stop);
            } else if (dist == 1) {
                memset(out, *from, length);
            } else {
                for (unsigned k = 0; k < length; ++k) out[k] = from[k];
            }
            out += length;
        }
    block_done:
        if (final_block) break;
    }
#undef REFILL
#undef NEED
#undef DROP
    return out == out_end ? 0 : -1;
}
