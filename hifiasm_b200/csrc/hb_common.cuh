// hb_common.cuh — shared device-side types and helpers of the overlap engine.
//
// Functions marked HB_HD are the bodies of one-thread-per-unit kernels; they
// compile as device code in the product (.cu files) and, ONLY inside
// tests/hostemu, as host code so their logic can be checked in the GPU-less
// build container.  The shipped library never runs them on the host.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "../../include/hifiasm_b200.h"

#ifdef __CUDACC__
#define HB_HD __host__ __device__ __forceinline__
#define HB_D __device__ __forceinline__
#define HB_HD_NI static __host__ __device__ __noinline__ /* large bodies with several call sites: one copy per kernel keeps the code in the instruction cache */
#else
#define HB_HD inline
#define HB_D inline
#define HB_HD_NI inline
struct ulonglong2 { unsigned long long x, y; };
#endif

// ---- read store in HBM: All_reads (Process_Read.h:115-146), flattened ------
// packed: 2-bit bases, byte = b0<<6|b1<<4|b2<<2|b3 (ha_compress_base,
// Process_Read.cpp:792); every read starts on a 32-byte boundary so kernels
// can stream whole sectors (128 bases per load).
struct DevReads {
	uint64_t n;
	const uint8_t *packed;
	const uint64_t *off;   // byte offset of read i (multiple of 8), n+1
	const uint32_t *len;   // read_length
	const uint64_t *noff;  // N_site offsets, n+1
	const uint32_t *npos;  // N positions (ascending per read)
};

// ---- high-count filter table: yak_ft_t (htab.cpp:1036) as open addressing --
struct DevFt {
	uint64_t mask;         // capacity-1 (0 => empty table)
	const uint64_t *key;
	const int32_t *val;    // 0 = empty slot; INT32_MAX for saturated counts
};

// ---- position index: ha_pt_t (htab.cpp:303-314) as ONE open-addressing table
// slot.x = minimizer hash, slot.y = offset<<12 | count (count>=2; 0 = empty);
// 16-byte slots => one 128-bit load per probe.
struct DevPt {
	uint64_t mask;
	const ulonglong2 *slot;
	const uint64_t *pos;   // ha_idxpos_t words, lists contiguous
};

#define HB_MZ_RID(i)  ((uint32_t)((i) & 0xfffffffULL))
#define HB_MZ_POS(i)  ((uint32_t)(((i) >> 28) & 0x7ffffffULL))
#define HB_MZ_REV(i)  ((uint32_t)(((i) >> 55) & 1ULL))
#define HB_MZ_SPAN(i) ((uint32_t)((i) >> 56))
#define HB_HIT_ID(h)  ((h).id_strand & 0x7fffffffu)
#define HB_HIT_ST(h)  ((h).id_strand >> 31)

HB_HD uint64_t hb_hash64(uint64_t key)
{ // yak_hash64_64, htab.h:150-160
	key = ~key + (key << 21);
	key = key ^ key >> 24;
	key = (key + (key << 3)) + (key << 8);
	key = key ^ key >> 14;
	key = (key + (key << 2)) + (key << 4);
	key = key ^ key >> 28;
	key = key + (key << 31);
	return key;
}

// bucket of a 64-bit minimizer hash in our tables (keys are already hashes; a
// multiplicative mix decorrelates the bucket from the 4096-way partition bits)
HB_HD uint64_t hb_bucket(uint64_t key, uint64_t mask) { return ((key * 0x9E3779B97F4A7C15ULL) >> 17) & mask; }

HB_HD int32_t hb_ft_lookup(const DevFt &ft, uint64_t y)
{ // ha_ft_cnt, htab.cpp:1064-1070
	if (ft.mask == 0) return 0;
	uint64_t i = hb_bucket(y, ft.mask);
	while (true) {
		int32_t v = ft.val[i];
		if (v == 0) return 0;
		if (ft.key[i] == y) return v;
		i = (i + 1) & ft.mask;
	}
}

// ha_pt_get, htab.cpp:518-527: returns count, *off = start in pos[]
HB_HD uint32_t hb_pt_lookup(const DevPt &pt, uint64_t hash, uint64_t *off)
{
	uint64_t i = hb_bucket(hash, pt.mask);
	while (true) {
#if defined(__CUDA_ARCH__)
		ulonglong2 s = __ldg(&pt.slot[i]);
#else
		ulonglong2 s = pt.slot[i];
#endif
		uint32_t c = (uint32_t)(s.y & 0xfffULL);
		if (c == 0) { *off = 0; return 0; }
		if (s.x == hash) { *off = s.y >> 12; return c; }
		i = (i + 1) & pt.mask;
	}
}

HB_HD int hb_base(const uint8_t *p, uint64_t i) { return (p[i >> 2] >> ((3 - (i & 3)) << 1)) & 3; }

// get_chainLen, Hash_Table.cpp:779-809
HB_HD int64_t hb_chain_len(int64_t xb, int64_t xe, int64_t xl, int64_t yb, int64_t ye, int64_t yl)
{
	int64_t xr, yr;
	if (xb <= yb) xb = 0; else xb -= yb;
	xr = xl - xe - 1; yr = yl - ye - 1;
	if (xr <= yr) xe = xl - 1; else xe += yr;
	return xe - xb + 1;
}

// in-place heapsort of bare 64-bit keys (where the reference radix-sorts key arrays without payload, any sort gives the same array)
HB_HD void hb_heapsort64(uint64_t *a, uint32_t n)
{
	if (n < 2) return;
	for (uint32_t s = n / 2; s-- > 0;) { uint32_t i = s; const uint64_t v = a[i]; for (;;) { uint32_t c = 2 * i + 1; if (c >= n) break; if (c + 1 < n && a[c + 1] > a[c]) c++; if (a[c] <= v) break; a[i] = a[c]; i = c; } a[i] = v; }
	for (uint32_t e = n - 1; e > 0; e--) { const uint64_t v = a[e]; a[e] = a[0]; uint32_t i = 0; for (;;) { uint32_t c = 2 * i + 1; if (c >= e) break; if (c + 1 < e && a[c + 1] > a[c]) c++; if (a[c] <= v) break; a[i] = a[c]; i = c; } a[i] = v; }
}
