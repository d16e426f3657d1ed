// hb_bloom.cuh — the blocked Bloom filter the reference puts in front of its k-mer counting (hifiasm -f, default 37): yak_bf_insert
// (htab.cpp:98-115) and the sub-table / block a k-mer falls into (ha_ct_insert_list 181-214, ha_ct_init 140-158).
// With the filter on, an occurrence is counted only if its 4 bits were all set before — so a k-mer's FIRST occurrence is lost unless it
// is a false positive, later ones always count, and the k-mer enters the table with 1 + 1.  The state of the filter at a k-mer's first
// occurrence is the union of the bits of the k-mers of the same 512-bit block that occurred earlier in (read, position) order, which is
// how the GPU evaluates it: distinct k-mers grouped by block, ordered by first occurrence, one thread walking a block (index.cu).
#pragma once
#include "hb_common.cuh"

// bf_local = bf_shift - 12 (bits of one sub-table's filter); active when bf_local >= 9 (yak_bf_init, htab.cpp:78-91)
HB_HD bool hb_bf_active(int bf_shift) { return bf_shift > 12 && bf_shift - 12 >= 9 && bf_shift - 12 + 9 <= 64; }
// the block of a k-mer among the 4096 << (bf_local - 9) blocks of all sub-tables
HB_HD uint64_t hb_bf_block(uint64_t hash, int bf_local) { const int xb = bf_local - 9; return (hash & 0xfff) << xb | ((hash >> 12) & ((1ULL << xb) - 1)); }
// yak_bf_insert on one block held as 8 words: sets the k-mer's 4 bits, returns how many were set before (4 = the occurrence counts)
HB_HD int hb_bf_insert(uint64_t st[8], uint64_t hash, int bf_local)
{
	const uint64_t x = hash >> 12; const int xb = bf_local - 9; int h2 = (int)(x >> bf_local & 511), z = (int)(x >> xb & 511), cnt = 0;
	if ((h2 & 31) == 0) h2 = (h2 + 1) & 511; // otherwise a few bits would be used repeatedly
	for (int i = 0; i < 4; i++, z = (z + h2) & 511) { const uint64_t u = 1ULL << (z & 63); cnt += (st[z >> 6] & u) != 0; st[z >> 6] |= u; }
	return cnt;
}
// count of a k-mer with `len` occurrences whose first one was (fp = 1) / was not (0) a false positive; 0 = never entered the table
HB_HD uint32_t hb_bf_count(uint64_t len, int fp) { const uint64_t ins = len - 1 + (uint64_t)fp; return ins == 0 ? 0u : (ins + 1 > 4095 ? 4095u : (uint32_t)(ins + 1)); }
