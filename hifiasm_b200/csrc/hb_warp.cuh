// hb_warp.cuh — the few warp collectives the warp-per-unit kernels use, behind names that also compile on the host.
//
// Kernel bodies that give one WARP to a unit (a read, an alignment) are written once: every lane calls the body with the same arguments and
// strides its loops by HB_WS.  In the product they are device code (HB_WS = 32, the collectives are shuffles / votes / shared-memory atomics);
// ONLY inside tests/hostemu they compile as host code with a warp of one lane, which checks the logic (not the synchronisation) against the
// golden vectors in the GPU-less container.
#pragma once
#include "hb_common.cuh"

#if defined(__CUDA_ARCH__)
#define HB_WS 32
HB_D int hb_lane() { return (int)(threadIdx.x & 31); }
HB_D void hb_wsync() { __syncwarp(); }
HB_D uint32_t hb_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
HB_D bool hb_any(bool p) { return __any_sync(0xffffffffu, p); }
HB_D uint32_t hb_bcast(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
HB_D uint64_t hb_bcast64(uint64_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
HB_D int32_t hb_wsum(int32_t v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
HB_D uint64_t hb_wsum64(uint64_t v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
HB_D uint32_t hb_wmax(uint32_t v) { for (int o = 16; o; o >>= 1) { const uint32_t u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; } return v; }
// exclusive prefix sum over the lanes; *total = sum over the warp
HB_D uint32_t hb_wscan(uint32_t v, uint32_t *total)
{
	uint32_t x = v;
	for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (hb_lane() >= o) x += y; }
	*total = __shfl_sync(0xffffffffu, x, 31);
	return x - v;
}
HB_D uint64_t hb_wscan64(uint64_t v, uint64_t *total)
{
	uint64_t x = v;
	for (int o = 1; o < 32; o <<= 1) { const uint64_t y = __shfl_up_sync(0xffffffffu, x, o); if (hb_lane() >= o) x += y; }
	*total = __shfl_sync(0xffffffffu, x, 31);
	return x - v;
}
HB_D void hb_atom_add64(uint64_t *p, uint64_t v) { atomicAdd((unsigned long long *)p, (unsigned long long)v); }
HB_D uint32_t hb_atom_add32(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
HB_D void hb_atom_or64(uint64_t *p, uint64_t v) { atomicOr((unsigned long long *)p, (unsigned long long)v); }
HB_D int hb_popc64(uint64_t v) { return __popcll(v); }
HB_D int hb_ctz64(uint64_t v) { return __ffsll((long long)v) - 1; } // v != 0
#else
#define HB_WS 1
inline int hb_lane() { return 0; }
inline void hb_wsync() {}
inline uint32_t hb_ballot(bool p) { return p ? 1u : 0u; }
inline bool hb_any(bool p) { return p; }
inline uint32_t hb_bcast(uint32_t v, int) { return v; }
inline uint64_t hb_bcast64(uint64_t v, int) { return v; }
inline int32_t hb_wsum(int32_t v) { return v; }
inline uint64_t hb_wsum64(uint64_t v) { return v; }
inline uint32_t hb_wmax(uint32_t v) { return v; }
inline uint32_t hb_wscan(uint32_t v, uint32_t *total) { *total = v; return 0; }
inline uint64_t hb_wscan64(uint64_t v, uint64_t *total) { *total = v; return 0; }
inline void hb_atom_add64(uint64_t *p, uint64_t v) { *p += v; }
inline uint32_t hb_atom_add32(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p += v; return o; }
inline void hb_atom_or64(uint64_t *p, uint64_t v) { *p |= v; }
inline int hb_popc64(uint64_t v) { return __builtin_popcountll(v); }
inline int hb_ctz64(uint64_t v) { return __builtin_ctzll(v); }
#endif
