// Worker pool + u8 / u16 -> i32 widening used by the host-buffer entry points (pss_expand.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <functional>

struct pss_pool;
pss_pool *pss_pool_create(int n_threads);
void pss_pool_destroy(pss_pool *p);
int pss_pool_size(const pss_pool *p);
size_t pss_pool_submit(pss_pool *p, std::function<void()> job);
void pss_pool_wait(pss_pool *p, size_t ticket);
void pss_widen_u16(const uint16_t *src, int32_t *dst, size_t n);
void pss_widen_u8(const uint8_t *src, int32_t *dst, size_t n);

// Row sink of pss_pileup_host_rows: the result chunks (8 / 16 / 32-bit columns in a pinned staging buffer) are
// packed straight into the caller's records, compacted to the emitted rows.
struct pss_row_pack {
    uint8_t *rows = nullptr;          // record 0
    int32_t row_bytes = 0;
    int n_want = 0;                   // value fields wanted
    int32_t f_idx[96], f_off[96], f_bytes[96];
    const uint8_t *chrom = nullptr;
    int32_t chrom_off = -1, chrom_bytes = 0, pos_off = -1, ref_off = -1, check_field = -1;
    int64_t pos_add = 0;
    const uint8_t *ref = nullptr;     // host reference bases of [ref_start, ref_start + ref_len)
    int64_t ref_start = 0, ref_len = 0;
    int64_t n_negative = 0;           // emitted values of check_field below zero (atomic adds)
};
int64_t pss_count_rows(const uint8_t *mask, int64_t n);
// positions [p0, p0 + len): mask[i] != 0 emits a row; field f of position p0 + i is element f * stride + i of src
void pss_pack_piece(pss_row_pack *K, const uint8_t *src, int64_t stride, int esz, const uint8_t *mask, int64_t p0,
                    int64_t len, int64_t row0);
