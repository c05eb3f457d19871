// Raw DEFLATE (RFC 1951) decoder for BGZF blocks: whole block in, whole block out, no streaming state.
// Replaces zlib's inflate() on the BAM decode path (what the reference gets from htslib's bgzf.c): a 64-bit
// bit buffer refilled once per symbol, an 11-bit primary litlen table (literals, lengths with their extra-bit
// counts and the end-of-block code in one lookup), an 8-bit primary offset table, subtables for longer codes,
// 8-byte match copies.  Anything unexpected -- a code set that is over-subscribed, a hole in an incomplete set, an
// offset before the start of the block, input or output running out -- returns -1 and the caller repeats the
// block with zlib, which then decides whether the block is corrupt.
#pragma once
#include <stddef.h>
#include <stdint.h>

// 0: dst[0, out_len) filled from a complete stream; -1: not decoded (fall back to zlib)
int pss_inflate_fast(const uint8_t *src, size_t in_len, uint8_t *dst, size_t out_len);
