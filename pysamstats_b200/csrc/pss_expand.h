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
