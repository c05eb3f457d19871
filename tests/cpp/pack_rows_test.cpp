// Host-side unit test of the row sink (pss_pack_piece / pss_count_rows, csrc/pss_expand.cpp): records packed from
// 8-bit result columns must equal a plain per-field gather, for every field count that takes the transposing path
// (8, >= 16), the scalar path, dense / gapped / sparse row masks; bytes before, between and after the records
// stay untouched (uncovered bytes inside a record may be zeroed).  Built and run by tests/test_host_logic.py.
#include "pss_expand.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
int main() {
    long total_bad = 0;
    for (int nf : {2, 6, 8, 16, 20, 24, 60}) for (int variant = 0; variant < 3; ++variant) {
        const int64_t stride = 5000, len = 4321, p0 = 777; const int esz = 1;
        std::vector<uint8_t> src((size_t)nf * stride);
        for (auto &b : src) b = rand();
        std::vector<uint8_t> mask(len, 1);
        if (variant == 1) for (int k = 0; k < 40; ++k) mask[rand() % len] = 0;
        if (variant == 2) for (auto &m : mask) m = rand() % 2;
        pss_row_pack K;
        const int base = 9, tail = variant == 0 ? 0 : 3;
        K.row_bytes = base + 4 * nf + tail; K.n_want = nf;
        for (int j = 0; j < nf; ++j) { K.f_idx[j] = j; K.f_off[j] = base + 4 * j; K.f_bytes[j] = 4; }
        const uint8_t chrom[4] = {'c', 'h', 'r', 0};
        K.chrom = chrom; K.chrom_off = 0; K.chrom_bytes = 4; K.pos_off = 4; K.ref_off = 8;
        std::vector<uint8_t> ref(6000, 'g'); K.ref = ref.data(); K.ref_start = 0; K.ref_len = 6000;
        int64_t cnt = pss_count_rows(mask.data(), len);
        std::vector<uint8_t> rows((size_t)(cnt + 2) * K.row_bytes, 0xEE);
        K.rows = rows.data();
        pss_pack_piece(&K, src.data(), stride, esz, mask.data(), p0, len, 1);
        int64_t row = 1, bad = 0;
        for (int64_t i = 0; i < len; ++i) if (mask[i]) {
            uint8_t *r = rows.data() + row * K.row_bytes;
            int32_t v; memcpy(&v, r + 4, 4); bad += v != p0 + i; bad += r[8] != 'G'; bad += memcmp(r, chrom, 4) != 0;
            for (int f = 0; f < nf; ++f) { memcpy(&v, r + base + 4 * f, 4); bad += v != src[(size_t)f * stride + i]; }
            for (int k = 0; k < tail; ++k) bad += r[base + 4 * nf + k] != 0xEE && r[base + 4 * nf + k] != 0;
            ++row;
        }
        for (int i = 0; i < K.row_bytes; ++i) { bad += rows[i] != 0xEE; bad += rows[(size_t)(cnt + 1) * K.row_bytes + i] != 0xEE; }
        printf("nf %d variant %d rows %lld bad %lld\n", nf, variant, (long long)cnt, (long long)bad);
        total_bad += bad;
    }
    return total_bad != 0;
}
