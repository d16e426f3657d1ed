// hb_kernels.cuh — the sm_100a kernels of the overlap / error-correction hot path (no tensor-core work exists on it: integer / byte streams; what
// matters is coalesced 128-bit access, shared-memory staging, warp collectives and grids sized from the SM count).
//
// The kernels are compiled in three translation units so that the build parallelises (HB_KERNELS_* selects the bodies; the argument structs and
// launch constants are visible everywhere):
//   HB_KERNELS_MAIN (engine.cu)   sketch, probe, expand, chain, chain post-filter, exact test, final merge, window pass, list emission, small helpers
//   HB_KERNELS_ECB  (kern_ecb.cu) step A (k_ec_overlap_fast / k_ec_overlap) and step B (k_ecb_prep, k_ecb_seg_fast, k_ecb_seg, k_ecb_seg_g, k_ecb_seg_w, k_ecb_merge)
//   HB_KERNELS_ECP  (kern_ecp.cu) phasing (k_ph_count, k_ph_decide) and window consensus (k_ec_cns_w)
// engine.cu launches the kernels of the other two units through the hb_k_* functions declared at the end of this file.
#pragma once
#include "hb_sketch.cuh"
#include "hb_final.cuh"
#include "hb_ecaln.cuh"
#include "hb_ecphase.cuh"
#include "hb_ecround.cuh"
#include "hb_eccns.cuh"
#include "hb_eccns_full.cuh"

#define HB_FULL 0xffffffffu
#include "hb_warp.cuh"

#ifdef HB_KERNELS_MAIN

// ----------------------------------------------------------------------------
// sketch
// ----------------------------------------------------------------------------

// ----------------------------------------------------------------------------
// two-stage sketch (production path; see hb_sketch.cuh)
//   k_sketch_events : one thread per read, straight-line scan -> ring-event stream
//   k_sketch_select : one block per read; tiles of events staged in shared memory,
//                     blocked prefix/suffix minima (van Herk) give the window
//                     minimum of every event, emissions are counted, scanned and
//                     written in order
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sketch_events(DevReads R, DevFt ft, SketchPar P, uint64_t r0, uint64_t nR, const uint64_t *__restrict__ ev_off,
                                                        ulonglong2 *evs, uint32_t *n_ev, uint32_t *tl)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nR) return;
	SkEv ev = { evs + ev_off[r] };
	uint32_t n = 0, t = 0;
	hb_sketch_events(R, ft, P, r0 + r, ev, &n, &t);
	n_ev[r] = n; tl[r] = t;
}

#define SK2_TS 1024
#define SK2_THREADS 128
#define SK2_EMPTY 0x7fffffffu
#define SK2_DUP 0x80000000u
struct SmemView64 { const uint64_t *p; __device__ __forceinline__ uint64_t operator[](int32_t i) const { return p[i]; } };
struct SmemView32 { const uint32_t *p; __device__ __forceinline__ uint32_t operator[](int32_t i) const { return p[i]; } };

// right-most argmin of an older range `o` and the adjacent newer range `nw` (tile-relative
// index | SK2_DUP when another entry of the range shares the minimum key; SK2_EMPTY = no entry)
static __device__ __forceinline__ uint32_t sk2_comb(uint32_t o, uint32_t nw, const uint64_t *sx, const uint64_t *sm)
{
	if ((o & ~SK2_DUP) == SK2_EMPTY) return nw;
	if ((nw & ~SK2_DUP) == SK2_EMPTY) return o;
	const uint32_t io = o & ~SK2_DUP, in = nw & ~SK2_DUP;
	const int c = sk_cmp(sx[in], sm[in], sx[io], sm[io]);
	return c < 0 ? nw : c > 0 ? o : (in | SK2_DUP);
}

__global__ void __launch_bounds__(SK2_THREADS) k_sketch_select(DevReads R, DevFt ft, SketchPar P, uint64_t r0, uint64_t nR, const uint64_t *__restrict__ ev_off,
                                                                const ulonglong2 *__restrict__ evs,
                                                                const uint32_t *__restrict__ n_ev, const uint32_t *__restrict__ tl_a,
                                                                const uint64_t *__restrict__ cap_off, hb_mz_t *mz, uint32_t *mz_l, uint32_t *mz_n, int *err)
{
	extern __shared__ uint64_t sk2_smem[];
	const int32_t w = P.w, k = P.k, tile = (SK2_TS / w) * w, capt = tile + w;
	uint64_t *sx = sk2_smem, *sm = sx + capt; uint32_t *sl = (uint32_t *)(sm + capt), *bufA = sl + capt, *bufB = bufA + capt, *bufC = bufB + capt, *bufD = bufC + capt, *cnt = bufD + capt;
	__shared__ uint32_t s_warp[SK2_THREADS / 32], s_on, s_ovf; __shared__ int32_t s_wmax[SK2_THREADS / 32], s_lastN;
	const uint64_t r = blockIdx.x;
	if (r >= nR) return;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int32_t T = (int32_t)n_ev[r];
	const uint64_t eb = ev_off[r];
	hb_mz_t *out = mz + cap_off[r]; uint32_t *out_l = mz_l + cap_off[r]; const uint32_t cap = (uint32_t)(cap_off[r + 1] - cap_off[r]);
	if (tid == 0) { s_on = 0; s_ovf = 0; s_lastN = -1; }
	__syncthreads();
	SmemView64 X = { sx }, M = { sm }; SmemView32 L = { sl };
	for (int32_t t0 = 0; t0 < T; t0 += tile) {
		const int32_t lo = t0 >= w ? t0 - w : 0, hi = T < t0 + tile ? T : t0 + tile, n = hi - lo;
		for (int32_t i = tid; i < n; i += SK2_THREADS) { const ulonglong2 e = __ldg(&evs[eb + lo + i]); sx[i] = e.x; sm[i] = e.y; bufA[i] = (uint32_t)i; bufC[i] = SK2_EMPTY; }
		__syncthreads();
		{ // l(t) = events since the last N base (0 on an N): block-wide running max of the N indices
			const int32_t pc = (n + SK2_THREADS - 1) / SK2_THREADS, c0 = tid * pc, c1 = c0 + pc < n ? c0 + pc : n;
			int32_t mx = -1;
			for (int32_t i = c0; i < c1; i++) if (sm[i] == SK_NMARK) mx = lo + i;
			int32_t inc = mx;
			for (int d = 1; d < 32; d <<= 1) { int32_t v = __shfl_up_sync(HB_FULL, inc, d); if (lane >= d && v > inc) inc = v; }
			if (lane == 31) s_wmax[wid] = inc;
			__syncthreads();
			int32_t run = s_lastN;
			for (int i = 0; i < wid; i++) if (s_wmax[i] > run) run = s_wmax[i];
			int32_t ex_ = __shfl_up_sync(HB_FULL, inc, 1); if (lane == 0) ex_ = -1;
			if (ex_ > run) run = ex_;
			for (int32_t i = c0; i < c1; i++) {
				if (sm[i] == SK_NMARK) { run = lo + i; sl[i] = 0; sm[i] = SK_DUMMY_META; }
				else sl[i] = (uint32_t)(lo + i - run);
			}
			// carry for the next tile: the last N index below its first staged event (t0 + tile - w)
			const int32_t nlo = t0 + tile - w - 1 - lo; // tile-relative index of the last event before the next tile's halo
			__syncthreads();
			if (nlo >= c0 && nlo < c1) s_lastN = (int32_t)(lo + nlo) - (int32_t)sl[nlo]; // an N event has l = 0, so this is its own index
		}
		__syncthreads();
		// window minima by doubling: level = argmin over the last 2^b events; acc picks the
		// levels named by the bits of w so that acc ends as the argmin over the last w events
		uint32_t *lev = bufA, *lev2 = bufB, *acc = bufC, *acc2 = bufD; int32_t acc_len = 0;
		for (int b = 0; (1 << b) <= w; b++) {
			const int32_t step = 1 << b; const bool take = (w >> b) & 1, grow = (2 << b) <= w;
			for (int32_t i = tid; i < n; i += SK2_THREADS) {
				if (take) acc2[i] = sk2_comb(i - acc_len >= 0 ? lev[i - acc_len] : SK2_EMPTY, acc[i], sx, sm);
				if (grow) lev2[i] = sk2_comb(i - step >= 0 ? lev[i - step] : SK2_EMPTY, lev[i], sx, sm);
			}
			__syncthreads();
			if (take) { uint32_t *q = acc; acc = acc2; acc2 = q; acc_len += step; }
			if (grow) { uint32_t *q = lev; lev = lev2; lev2 = q; }
		}
		// acc[i] = right-most argmin (tile-relative) of events (t-w, t], t = lo + i
		const int32_t per = (tile + SK2_THREADS - 1) / SK2_THREADS, a0 = t0 + tid * per, a1 = (a0 + per < hi) ? a0 + per : hi;
		uint32_t mysum = 0;
		for (int32_t t = a0; t < a1; t++) {
			const uint32_t vp = t > 0 ? acc[t - 1 - lo] : SK2_EMPTY, vc = acc[t - lo];
			const int32_t mp = t > 0 ? lo + (int32_t)(vp & ~SK2_DUP) : -1, mc = lo + (int32_t)(vc & ~SK2_DUP);
			const uint32_t c = sk2_emit(X, M, L, lo, t, mp, mc, w, k, (hb_mz_t *)0, (uint32_t *)0, (vp & SK2_DUP) != 0, (vc & SK2_DUP) != 0);
			cnt[t - t0] = c; mysum += c;
		}
		uint32_t inc = mysum;
		for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(HB_FULL, inc, d); if (lane >= d) inc += v; }
		if (lane == 31) s_warp[wid] = inc;
		__syncthreads();
		uint32_t wbase = 0, total = 0;
		for (int i = 0; i < SK2_THREADS / 32; i++) { if (i < wid) wbase += s_warp[i]; total += s_warp[i]; }
		uint32_t o = s_on + wbase + inc - mysum;
		const bool last = hi == T;
		const uint32_t vlast = last ? acc[T - 1 - lo] : SK2_EMPTY;
		const bool flush = last && sx[vlast & ~SK2_DUP] != ~0ULL;
		const bool fits = s_on + total + (flush ? 1u : 0u) <= cap;
		if (fits) {
			for (int32_t t = a0; t < a1; t++) {
				const uint32_t c = cnt[t - t0];
				if (c) {
					const uint32_t vp = t > 0 ? acc[t - 1 - lo] : SK2_EMPTY, vc = acc[t - lo];
					const int32_t mp = t > 0 ? lo + (int32_t)(vp & ~SK2_DUP) : -1, mc = lo + (int32_t)(vc & ~SK2_DUP);
					sk2_emit(X, M, L, lo, t, mp, mc, w, k, out + o, out_l + o, (vp & SK2_DUP) != 0, (vc & SK2_DUP) != 0);
					o += c;
				}
			}
			if (flush && tid == 0) { const uint32_t m = vlast & ~SK2_DUP, at = s_on + total; out[at].x = sx[m]; out[at].info = sm[m]; out_l[at] = sl[m]; } // sketch.cpp:571-573
		}
		__syncthreads();
		if (tid == 0) { if (fits) s_on += total + (flush ? 1u : 0u); else s_ovf = 1; }
		__syncthreads();
		if (s_ovf) break;
	}
	if (s_ovf) { if (tid == 0) { mz_n[r] = 0; atomicOr(err, 1); } return; }
	if (tid == 0) {
		SketchOut o; o.mz = out; o.l = out_l; o.cap = cap; o.n = s_on; o.ovf = 0;
		if (P.sample_dist > w && ft.mask != 0) sk_select_mz_h(o, (int32_t)R.len[r0 + r], P.sample_dist, P.rewin, k, (int32_t)tl_a[r]); // sketch.cpp:575
		s_on = o.n; mz_n[r] = o.n;
	}
	__syncthreads();
	const uint32_t rid_out = (uint32_t)(r0 + r) & 0xfffffff;
	for (uint32_t i = tid; i < s_on; i += SK2_THREADS) out[i].info = (out[i].info & ~0xfffffffULL) | rid_out; // sketch.cpp:577
}

// strided per-read slices -> dense array (warp per read)
__global__ void k_compact_mz(uint64_t nR, const uint64_t *__restrict__ cap_off, const uint64_t *__restrict__ off, const hb_mz_t *__restrict__ in, hb_mz_t *__restrict__ out)
{
	uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= nR) return;
	const uint4 *s = (const uint4 *)(in + cap_off[w]); uint4 *d = (uint4 *)(out + off[w]);
	uint32_t n = (uint32_t)(off[w + 1] - off[w]);
	for (uint32_t i = hb_lane(); i < n; i += 32) d[i] = s[i];
}

// ----------------------------------------------------------------------------
// probe: ha_pt_get for every query minimizer (htab.cpp:518; anchor.cpp:1011-1016)
// seeds[i] = offset<<12 | n ; spre[i] = anchors of the read before minimizer i
// ----------------------------------------------------------------------------
__global__ void k_probe_count(DevPt pt, uint64_t nR, const uint64_t *__restrict__ mz_off, const hb_mz_t *__restrict__ mz,
                              uint64_t *__restrict__ seeds, uint32_t *__restrict__ spre, uint32_t *__restrict__ a_cnt)
{
	uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= nR) return;
	int lane = hb_lane();
	uint64_t b = mz_off[w], e = mz_off[w + 1]; uint32_t run = 0;
	for (uint64_t i = b; i < e; i += 32) {
		uint64_t s = i + lane, off = 0; uint32_t c = 0;
		if (s < e) c = hb_pt_lookup(pt, mz[s].x, &off);
		uint32_t inc = c;
		for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(HB_FULL, inc, d); if (lane >= d) inc += v; }
		if (s < e) { seeds[s] = off << 12 | c; spre[s] = run + inc - c; }
		run += __shfl_sync(HB_FULL, inc, 31);
	}
	if (lane == 0) a_cnt[w] = run;
}

// ----------------------------------------------------------------------------
// expand: the seed-hash probe kernel proper.  Half-warp per query minimizer:
// 16 lanes x 128-bit loads walk the minimizer's ha_idxpos_t list; every hit
// becomes a k_mer_hit (anchor.cpp:1023-1037 + 1055-1076) written with 128-bit
// stores at its final place in the read's (unsorted) anchor run.
// Algorithmic bytes: 16 (minimizer) + 8 (seed) + 4 (prefix) per minimizer,
// 8 in + 16 out per hit.
// ----------------------------------------------------------------------------
template <int EXP_U, int MINB> // EXP_U query minimizers in flight per half-warp
__global__ void __launch_bounds__(256, MINB) k_expand(DevReads R, DevPt pt, uint64_t r0, uint64_t n_mz, const hb_mz_t *__restrict__ mz, const uint64_t *__restrict__ seeds,
                         const uint32_t *__restrict__ spre, const uint64_t *__restrict__ a_off, uint64_t a_base, const uint32_t *__restrict__ w_tab, hb_hit_t *__restrict__ raw)
{ // mz/seeds/spre: the batch's minimizers; a_off: anchor offsets indexed by (read - r0); a_base = a_off of the batch's first read
	const uint64_t hw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, n_hw = ((uint64_t)gridDim.x * blockDim.x) >> 4;
	const int l16 = threadIdx.x & 15;
	for (uint64_t s0 = hw * EXP_U; s0 < n_mz; s0 += n_hw * EXP_U) {
		uint64_t sd[EXP_U], zi[EXP_U]; uint32_t sp[EXP_U];
		// stage 1: all independent loads of EXP_U minimizers are issued before any is used
#pragma unroll
		for (int u = 0; u < EXP_U; u++) {
			const uint64_t s = s0 + u; const bool in = s < n_mz;
			sd[u] = in ? __ldg(&seeds[s]) : 0; zi[u] = in ? __ldg(&mz[s].info) : 0; sp[u] = in ? __ldg(&spre[s]) : 0;
		}
		uint64_t ao[EXP_U]; ulonglong2 v[EXP_U];
#pragma unroll
		for (int u = 0; u < EXP_U; u++) {
			const uint32_t n = (uint32_t)(sd[u] & 0xfff); const uint64_t off = sd[u] >> 12, a0 = off & ~1ULL;
			ao[u] = n ? __ldg(&a_off[HB_MZ_RID(zi[u]) - r0]) : 0;
			v[u].x = v[u].y = 0;
			if (n && (uint32_t)(2 * l16) < (uint32_t)(off - a0) + n) v[u] = __ldg((const ulonglong2 *)(pt.pos + a0 + 2 * l16));
		}
		// stage 2: first 32 list slots of each minimizer (covers the typical list), target-length gathers batched
		uint32_t tl[EXP_U][2];
#pragma unroll
		for (int u = 0; u < EXP_U; u++) {
			const uint32_t n = (uint32_t)(sd[u] & 0xfff), head = (uint32_t)((sd[u] >> 12) & 1);
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const int32_t j = (int32_t)(2 * l16 + h) - (int32_t)head;
				tl[u][h] = (j >= 0 && j < (int32_t)n) ? __ldg(&R.len[HB_MZ_RID(h ? v[u].y : v[u].x)]) : 0;
			}
		}
#pragma unroll
		for (int u = 0; u < EXP_U; u++) {
			const uint32_t n = (uint32_t)(sd[u] & 0xfff);
			if (n == 0) continue;
			const uint64_t off = sd[u] >> 12, a0 = off & ~1ULL; const uint32_t head = (uint32_t)(off - a0), span = head + n;
			const uint32_t zrev = HB_MZ_REV(zi[u]), zpos = HB_MZ_POS(zi[u]), zspan = HB_MZ_SPAN(zi[u]);
			const uint32_t cnt = __ldg(&w_tab[n]) << 8 | (zspan <= 255 ? zspan : 255);
			hb_hit_t *dst = raw + (ao[u] - a_base) + sp[u];
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const int32_t j = (int32_t)(2 * l16 + h) - (int32_t)head;
				if (j < 0 || j >= (int32_t)n) continue;
				const uint64_t y = h ? v[u].y : v[u].x; const uint32_t rev = zrev ^ HB_MZ_REV(y);
				uint4 o; o.x = HB_MZ_RID(y) | rev << 31; o.y = rev ? tl[u][h] - 1 - (HB_MZ_POS(y) + 1 - HB_MZ_SPAN(y)) : HB_MZ_POS(y); o.z = zpos; o.w = cnt;
				*(uint4 *)(dst + j) = o;
			}
			for (uint32_t e = 32 + 2 * l16; e < span; e += 32) { // long lists: the rest, 32 slots per step
				const ulonglong2 w = __ldg((const ulonglong2 *)(pt.pos + a0 + e));
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const int32_t j = (int32_t)(e + h) - (int32_t)head;
					if (j >= (int32_t)n) continue;
					const uint64_t y = h ? w.y : w.x; const uint32_t tid = HB_MZ_RID(y), rev = zrev ^ HB_MZ_REV(y), t2 = __ldg(&R.len[tid]);
					uint4 o; o.x = tid | rev << 31; o.y = rev ? t2 - 1 - (HB_MZ_POS(y) + 1 - HB_MZ_SPAN(y)) : HB_MZ_POS(y); o.z = zpos; o.w = cnt;
					*(uint4 *)(dst + j) = o;
				}
			}
		}
	}
}

#define GRP_EMPTY 0xffffffffu // GroupDir.slot of a group that is not chained (the read's hits on itself); groups are built in group.cu

// ----------------------------------------------------------------------------
// chain: one thread per (query, target) group
// ----------------------------------------------------------------------------
struct ChainArgs {
	DevReads R; uint64_t r0; const GroupDir *dir; const uint32_t *dir_n; const uint64_t *a_off; uint64_t a_base; const uint64_t *c_off;
	hb_hit_t *hits, *chits; int32_t *f, *p, *ii; int64_t *t; hb_chain_t *ch; uint32_t *slot_read; uint64_t *fc; ChainPar P; int *err; unsigned long long *dbg; uint32_t *work;
};

// ----------------------------------------------------------------------------
// chain, warp per group.  The common case — every strand block of the group is
// already co-linear, which is what quick_ck_lchain (Hash_Table.cpp:2007-2094)
// detects with a linear scan — is evaluated by the whole warp: 32 link scores
// per step, chain scores by warp prefix sums, best end / minimum by warp
// reductions, chain anchors copied with 128-bit loads.  A group that needs the DP runs
// warp_chain_dp (the predecessor scan of an anchor spread over the lanes, the reference's
// early-exit order kept); only the ordering of an unordered group (rare) is left to lane 0.
// ----------------------------------------------------------------------------
static __device__ __forceinline__ hb_hit_t hit_shfl_up1(const hb_hit_t &h)
{
	hb_hit_t r;
	r.id_strand = __shfl_up_sync(HB_FULL, h.id_strand, 1); r.offset = __shfl_up_sync(HB_FULL, h.offset, 1);
	r.self_offset = __shfl_up_sync(HB_FULL, h.self_offset, 1); r.cnt = __shfl_up_sync(HB_FULL, h.cnt, 1);
	return r;
}
static __device__ __forceinline__ hb_hit_t hit_ld(const hb_hit_t *p) { uint4 v = *(const uint4 *)p; hb_hit_t h; h.id_strand = v.x; h.offset = v.y; h.self_offset = v.z; h.cnt = v.w; return h; }
static __device__ __forceinline__ uint64_t hit_okey(const hb_hit_t &h) { return (uint64_t)(h.id_strand >> 31) << 63 | (uint64_t)h.self_offset << 32 | h.offset; }

// is a[0..n) ordered by (strand, self_offset, offset)?  (warp-uniform result); *kb = first index of strand 1 (n if none)
static __device__ bool warp_group_ordered(const hb_hit_t *a, int32_t n, int lane, int32_t *kb)
{
	bool bad = false; int32_t first1 = n;
	for (int32_t base = 0; base < n; base += 32) {
		int32_t z = base + lane; bool valid = z < n;
		hb_hit_t h = hit_ld(a + (valid ? z : n - 1));
		uint64_t kz = hit_okey(h), kp = __shfl_up_sync(HB_FULL, kz, 1);
		if (lane == 0) kp = base ? hit_okey(hit_ld(a + base - 1)) : 0;
		if (valid && z > 0 && kp > kz) bad = true;
		unsigned m1 = __ballot_sync(HB_FULL, valid && (h.id_strand >> 31));
		if (m1 && first1 == n) first1 = base + __ffs(m1) - 1;
	}
	*kb = first1;
	return !__any_sync(HB_FULL, bad);
}

struct QBlock { bool resolved; int64_t msc0, msc_i0, plus0; };
// quick_ck_lchain for one strand block [l,k) (Hash_Table.cpp:2023-2086), warp-parallel
static __device__ QBlock warp_quick_block(const hb_hit_t *a, int32_t l, int32_t k, const ChainPar &P, int64_t xl, int64_t yl, int lane, int32_t *f = 0, int32_t *p = 0)
{
	QBlock q; q.resolved = false; q.msc0 = q.msc_i0 = q.plus0 = 0;
	hb_hit_t first = hit_ld(a + l);
	int64_t carry_f = first.cnt & 0xffu, msc0 = carry_f, msc_i0 = l, plus0 = 0, ddt = 0;
	if (f && lane == 0) { f[l] = (int32_t)carry_f; p[l] = -1; }
	for (int32_t base = l + 1; base < k; base += 32) {
		int32_t z = base + lane; bool valid = z < k;
		hb_hit_t h = hit_ld(a + (valid ? z : k - 1)), hp = hit_shfl_up1(h);
		if (lane == 0) hp = hit_ld(a + base - 1);
		int32_t dd = 0, s = HB_LINK_FAIL; bool viol = false;
		if (valid) {
			if (h.self_offset <= hp.self_offset || h.offset <= hp.offset) viol = true; // is_srt, Hash_Table.cpp:2090
			else { s = hb_link_sc(h, hp, P, xl, yl, &dd); if (s == HB_LINK_FAIL) viol = true; }
		}
		int64_t inc = (valid && !viol) ? (int64_t)s : 0;
		for (int d = 1; d < 32; d <<= 1) { int64_t v = __shfl_up_sync(HB_FULL, inc, d); if (lane >= d) inc += v; }
		int64_t fz = carry_f + inc;
		if (valid && !viol && fz < (int64_t)(h.cnt & 0xffu)) viol = true; // sc < csc, Hash_Table.cpp:2059
		if (__any_sync(HB_FULL, viol)) return q; // the scan would stop before the block's end: not resolved
		if (f && valid) { f[z] = (int32_t)fz; p[z] = z - 1; }
		// running best end (>=: the later anchor wins ties), minimum, gap sum
		int64_t bf = valid ? fz : INT64_MIN, bi = valid ? z : -1, mf = valid ? fz : INT64_MAX, sd = valid ? dd : 0;
		for (int d = 16; d; d >>= 1) {
			int64_t of = __shfl_xor_sync(HB_FULL, bf, d), oi = __shfl_xor_sync(HB_FULL, bi, d), om = __shfl_xor_sync(HB_FULL, mf, d), od = __shfl_xor_sync(HB_FULL, sd, d);
			if (of > bf || (of == bf && oi > bi)) { bf = of; bi = oi; }
			if (om < mf) mf = om;
			sd += od;
		}
		if (bf >= msc0) { msc0 = bf; msc_i0 = bi; }
		if (mf < plus0) plus0 = mf;
		ddt += sd;
		int last = (k - base < 32 ? k - base : 32) - 1;
		carry_f = __shfl_sync(HB_FULL, fz, last);
	}
	if (msc_i0 != k - 1) return q;
	if (k - l >= 2 && ddt > 16) { // Hash_Table.cpp:2071
		hb_hit_t e = hit_ld(a + k - 1);
		if (ddt > hb_link_bw(e, first, P.bw_rate, xl, yl)) return q;
	}
	q.resolved = true; q.msc0 = msc0; q.msc_i0 = msc_i0; q.plus0 = plus0;
	return q;
}

// One pass over an (expected ordered, expected co-linear) group: order check, strand-block
// boundary, quick_ck_lchain for both strand blocks (segmented warp prefix sums), 32 anchors per
// step with the next 32 already in flight.  ok = false sends the group to the general path.
struct QFused { bool ok; int32_t kb; QBlock b[2]; };
static __device__ QFused warp_quick_fused(const hb_hit_t *a, int32_t n, const ChainPar &P, int64_t xl, int64_t yl, int lane)
{
	QFused R; R.ok = false; R.kb = n; R.b[0].resolved = R.b[1].resolved = false; R.b[0].msc0 = R.b[0].msc_i0 = R.b[0].plus0 = R.b[1].msc0 = R.b[1].msc_i0 = R.b[1].plus0 = 0;
	int64_t carry_f = 0, msc0[2] = { INT64_MIN, INT64_MIN }, msc_i0[2] = { -1, -1 }, plus0[2] = { 0, 0 }, ddt[2] = { 0, 0 };
	hb_hit_t first[2], lasth, prev_last; first[0].id_strand = first[1].id_strand = 0; first[0].offset = first[0].self_offset = first[0].cnt = first[1].offset = first[1].self_offset = first[1].cnt = 0;
	prev_last = first[0]; lasth = first[0];
	int32_t cntb[2] = { 0, 0 };
	hb_hit_t nxt = hit_ld(a + (lane < n ? lane : n - 1));
	for (int32_t base = 0; base < n; base += 32) {
		const int32_t z = base + lane; const bool valid = z < n;
		const hb_hit_t h = nxt;
		if (base + 32 < n) nxt = hit_ld(a + (z + 32 < n ? z + 32 : n - 1)); // prefetch
		hb_hit_t hp = hit_shfl_up1(h); if (lane == 0) hp = prev_last;
		const uint32_t sb = h.id_strand >> 31;
		const bool start = valid && (z == 0 || sb != (hp.id_strand >> 31));
		bool bad = false, viol = false; int32_t dd = 0, s = 0;
		if (valid && z > 0 && hit_okey(hp) > hit_okey(h)) bad = true; // not in (strand, self_offset, offset) order
		if (valid && !start) {
			if (h.self_offset <= hp.self_offset || h.offset <= hp.offset) viol = true;
			else { s = hb_link_sc(h, hp, P, xl, yl, &dd); if (s == HB_LINK_FAIL) viol = true; }
		}
		if (__any_sync(HB_FULL, bad || viol)) return R;
		const int64_t v = !valid ? 0 : start ? (int64_t)(h.cnt & 0xffu) : (int64_t)s;
		int64_t inc = v;
		for (int d = 1; d < 32; d <<= 1) { int64_t u = __shfl_up_sync(HB_FULL, inc, d); if (lane >= d) inc += u; }
		const unsigned sm = __ballot_sync(HB_FULL, start);
		const unsigned below = sm & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1));
		const int sl = below ? 31 - __clz(below) : -1;
		const int64_t excl_at_start = __shfl_sync(HB_FULL, inc - v, sl < 0 ? 0 : sl);
		const int64_t fz = sl < 0 ? carry_f + inc : inc - excl_at_start;
		if (__any_sync(HB_FULL, valid && !start && fz < (int64_t)(h.cnt & 0xffu))) return R; // sc < csc, Hash_Table.cpp:2059
		if (sm) { // a block starts in this chunk
			const int fl = __ffs(sm) - 1; // there is at most one start per strand; strand of the starting lane
			for (unsigned m = sm; m; m &= m - 1) {
				const int l0 = __ffs(m) - 1; const uint32_t bs = __shfl_sync(HB_FULL, sb, l0);
				first[bs].id_strand = __shfl_sync(HB_FULL, h.id_strand, l0); first[bs].offset = __shfl_sync(HB_FULL, h.offset, l0);
				first[bs].self_offset = __shfl_sync(HB_FULL, h.self_offset, l0); first[bs].cnt = __shfl_sync(HB_FULL, h.cnt, l0);
				if (bs == 1) R.kb = base + l0;
			}
			(void)fl;
		}
		// per-block running best end (>=: later wins), minimum, gap sum, count
		const unsigned vm = __ballot_sync(HB_FULL, valid), m1 = __ballot_sync(HB_FULL, valid && sb), m0 = vm & ~m1;
		for (int bs = 0; bs < 2; bs++) {
			const unsigned mm = bs ? m1 : m0;
			if (!mm) continue;
			const bool in = (mm >> lane) & 1;
			int64_t bf = in ? fz : INT64_MIN, bi = in ? z : -1, mf = in ? fz : INT64_MAX, sd = (in && !start) ? dd : 0;
			for (int d = 16; d; d >>= 1) {
				int64_t of = __shfl_xor_sync(HB_FULL, bf, d), oi = __shfl_xor_sync(HB_FULL, bi, d), om = __shfl_xor_sync(HB_FULL, mf, d), od = __shfl_xor_sync(HB_FULL, sd, d);
				if (of > bf || (of == bf && oi > bi)) { bf = of; bi = oi; }
				if (om < mf) mf = om;
				sd += od;
			}
			if (bf >= msc0[bs]) { msc0[bs] = bf; msc_i0[bs] = bi; }
			if (mf < plus0[bs]) plus0[bs] = mf;
			ddt[bs] += sd; cntb[bs] += __popc(mm);
		}
		const int lastl = (n - base < 32 ? n - base : 32) - 1;
		carry_f = __shfl_sync(HB_FULL, fz, lastl);
		prev_last.id_strand = __shfl_sync(HB_FULL, h.id_strand, lastl); prev_last.offset = __shfl_sync(HB_FULL, h.offset, lastl);
		prev_last.self_offset = __shfl_sync(HB_FULL, h.self_offset, lastl); prev_last.cnt = __shfl_sync(HB_FULL, h.cnt, lastl);
		if (m0) { const int l0 = 31 - __clz(m0); lasth.id_strand = __shfl_sync(HB_FULL, h.id_strand, l0); lasth.offset = __shfl_sync(HB_FULL, h.offset, l0); lasth.self_offset = __shfl_sync(HB_FULL, h.self_offset, l0); lasth.cnt = __shfl_sync(HB_FULL, h.cnt, l0); }
	}
	// lasth = last anchor of block 0 (if any); prev_last = last anchor of the group (= last of block 1 when it exists)
	for (int bs = 0; bs < 2; bs++) {
		if (!cntb[bs]) continue;
		const int32_t l = bs ? R.kb : 0, kk = bs ? n : R.kb;
		if (msc_i0[bs] != kk - 1) return R;
		if (kk - l >= 2 && ddt[bs] > 16) { // Hash_Table.cpp:2071
			const hb_hit_t e = (bs == 0 && cntb[1]) ? lasth : prev_last;
			if (ddt[bs] > hb_link_bw(e, first[bs], P.bw_rate, xl, yl)) return R;
		}
		R.b[bs].resolved = true; R.b[bs].msc0 = msc0[bs]; R.b[bs].msc_i0 = msc_i0[bs]; R.b[bs].plus0 = plus0[bs];
	}
	R.ok = true;
	return R;
}

// The chaining DP of lchain_qdp_mcopy_fast (Hash_Table.cpp:2124-2176) with the
// predecessor scan spread over the warp: 32 predecessors per step, link scores in
// parallel, the sequential semantics (running best, n_skip early stop, t[] marks)
// resolved with a prefix max + ballots + a 32-step scalar walk.  Marks are written
// optimistically by every lane of a step: a lane's own mark test can only be
// affected by lanes that precede it in scan order (p[j'] < j'), and marks left
// behind past a break carry the value i, which no later test looks for.
static __device__ void warp_chain_dp(const hb_hit_t *a, int32_t kb, int32_t *f, int32_t *p, int64_t *t, int32_t *ii, const ChainPar &P, int64_t xl, int64_t yl, ChainState &S, int lane)
{
	int64_t msc = S.msc, msc_i = S.msc_i, movl = S.movl, plus = S.plus, st = S.si, max_ii = -1;
	for (int64_t i = S.si; i < S.ei; ++i) {
		const hb_hit_t ai = hit_ld(a + i); const uint32_t si_ = ai.id_strand >> 31;
		int64_t max_f = ai.cnt & 0xffu, max_j = -1, end_j = -1; int32_t n_skip = 0; bool broke = false;
		if (i - st > P.max_iter) st = i - P.max_iter;
		if (si_ && st < kb) st = kb; // while (a[i].strand != a[st].strand) ++st;
		for (int64_t jtop = i - 1; jtop >= st && !broke; jtop -= 32) {
			const int64_t j = jtop - lane; const bool valid = j >= st;
			int32_t s = HB_LINK_FAIL; int64_t cand = INT64_MIN; int32_t pj = -1;
			if (valid) { hb_hit_t aj = hit_ld(a + j); s = hb_link_sc(ai, aj, P, xl, yl, 0); }
			const bool proc = valid && s != HB_LINK_FAIL;
			if (proc) { cand = (int64_t)s + f[j]; pj = p[j]; if (pj >= 0) t[pj] = i; }
			__syncwarp();
			const bool hit = proc && t[j] == (int64_t)(int32_t)i;
			int64_t pm = cand;
			for (int d = 1; d < 32; d <<= 1) { int64_t v = __shfl_up_sync(HB_FULL, pm, d); if (lane >= d && v > pm) pm = v; }
			int64_t excl = __shfl_up_sync(HB_FULL, pm, 1); if (lane == 0) excl = INT64_MIN;
			const int64_t run = excl > max_f ? excl : max_f;
			const bool imp = proc && cand > run;
			const unsigned impm = __ballot_sync(HB_FULL, imp), hitm = __ballot_sync(HB_FULL, hit && !imp), procm = __ballot_sync(HB_FULL, proc);
			int brk = -1;
			for (unsigned m = procm; m; m &= m - 1) {
				int b = __ffs(m) - 1;
				if (impm >> b & 1) { if (n_skip > 0) --n_skip; }
				else if (hitm >> b & 1) { if (++n_skip > P.max_skip) { brk = b; break; } }
			}
			const int lim = brk < 0 ? 32 : brk;
			int64_t bf = (imp && lane < lim) ? cand : INT64_MIN; int bl = lane;
			for (int d = 16; d; d >>= 1) {
				int64_t of = __shfl_xor_sync(HB_FULL, bf, d); int ol = __shfl_xor_sync(HB_FULL, bl, d);
				if (of > bf || (of == bf && ol < bl)) { bf = of; bl = ol; }
			}
			if (bf > max_f) { max_f = bf; max_j = jtop - bl; }
			if (brk >= 0) { broke = true; end_j = jtop - brk; }
		}
		if (!broke) end_j = st - 1;
		bool redo = max_ii < 0;
		hb_hit_t am; am.id_strand = am.offset = am.self_offset = am.cnt = 0;
		if (!redo) { am = hit_ld(a + max_ii); redo = (int64_t)ai.self_offset > (int64_t)am.self_offset + P.max_dis || si_ != (am.id_strand >> 31); }
		if (redo) { // Hash_Table.cpp:2145-2152
			int64_t bf = INT64_MIN, bj = -1; bool stop = false;
			for (int64_t jtop = i - 1; jtop >= st && !stop; jtop -= 32) {
				const int64_t j = jtop - lane; bool ok = j >= st;
				if (ok) { hb_hit_t aj = hit_ld(a + j); ok = (int64_t)ai.self_offset <= (int64_t)P.max_dis + (int64_t)aj.self_offset && si_ == (aj.id_strand >> 31); }
				const unsigned okm = __ballot_sync(HB_FULL, ok);
				// the scan stops at the first predecessor that fails the test: only the leading run of ok lanes counts
				const unsigned lead = okm == HB_FULL ? HB_FULL : ((1u << (__ffs(~okm) - 1)) - 1);
				if (lead != HB_FULL) stop = true;
				int64_t cf = (lead >> lane & 1) ? (int64_t)f[j] : INT64_MIN, cj = j;
				for (int d = 16; d; d >>= 1) {
					int64_t of = __shfl_xor_sync(HB_FULL, cf, d), oj = __shfl_xor_sync(HB_FULL, cj, d);
					if (of > cf || (of == cf && oj > cj)) { cf = of; cj = oj; }
				}
				if (cf > bf) { bf = cf; bj = cj; } // strict: an earlier (larger j) maximum is kept
			}
			max_ii = bf > (int64_t)INT32_MIN ? bj : -1;
			if (max_ii >= 0) am = hit_ld(a + max_ii);
		}
		if (max_ii >= 0 && max_ii < end_j && si_ == (am.id_strand >> 31)) {
			int32_t tmp = hb_link_sc(ai, am, P, xl, yl, 0);
			if (tmp != HB_LINK_FAIL && max_f < (int64_t)tmp + f[max_ii]) { max_f = (int64_t)tmp + f[max_ii]; max_j = max_ii; }
		}
		if (lane == 0) { f[i] = (int32_t)max_f; p[i] = (int32_t)max_j; ii[i] = 0; }
		__syncwarp();
		if (max_ii < 0 || ((int64_t)ai.self_offset <= (int64_t)P.max_dis + (int64_t)am.self_offset && si_ == (am.id_strand >> 31) && (int64_t)f[max_ii] < max_f)) max_ii = i;
		if (max_f >= msc) {
			int64_t ovl = hb_chain_len(ai.self_offset, ai.self_offset, xl, ai.offset, ai.offset, yl);
			if (max_f > msc || ovl < movl) { msc = max_f; msc_i = i; movl = ovl; }
		}
		if (max_f < plus) plus = max_f;
	}
	S.plus = plus; S.msc = msc; S.msc_i = msc_i; S.movl = movl;
}

__global__ void __launch_bounds__(128) k_chain_warp(ChainArgs A)
{
	const uint32_t n = *A.dir_n; const int lane = hb_lane();
	for (;;) { // groups are handed out one at a time: a few of them (those needing the DP) cost 100x the rest
		uint32_t g = 0;
		if (lane == 0) g = atomicAdd(A.work, 1u);
		g = __shfl_sync(HB_FULL, g, 0);
		if (g >= n) break;
		GroupDir d = A.dir[g];
		const uint64_t ab = A.a_off[d.read] - A.a_base + d.start; hb_hit_t *a = A.hits + ab; const int32_t an = (int32_t)d.count;
		int32_t kb = an; bool ordered = true;
		if (d.slot == GRP_EMPTY) { ordered = warp_group_ordered(a, an, lane, &kb); if (!ordered && lane == 0) hb_order_group(a, an); __syncwarp(); continue; }
		const uint64_t cb = A.c_off[d.read] + d.slot;
		const int32_t ns = d.count >= (uint32_t)A.P.mcopy_khit_cutoff ? A.P.mcopy_num : 1;
		const int64_t xl = A.R.len[A.r0 + d.read], yl = A.R.len[HB_HIT_ID(hit_ld(a))];
		bool simple; int64_t msc = INT32_MIN, msc_i = INT32_MIN, movl = INT32_MAX, plus = 0, other_max = INT64_MIN; int32_t lb = 0;
		{
			const QFused Q = warp_quick_fused(a, an, A.P, xl, yl, lane);
			simple = Q.ok; kb = Q.kb;
			for (int b = 0; b < 2 && simple; b++) { // combine the strand blocks, Hash_Table.cpp:2072-2085
				if (!Q.b[b].resolved) continue;
				const QBlock &q = Q.b[b]; bool took = false;
				if (q.msc0 >= msc) {
					hb_hit_t e = hit_ld(a + q.msc_i0);
					int64_t movl0 = hb_chain_len(e.self_offset, e.self_offset, xl, e.offset, e.offset, yl);
					if (q.msc0 > msc || movl0 < movl) { if (msc_i != INT32_MIN) other_max = msc; msc = q.msc0; msc_i = q.msc_i0; movl = movl0; lb = b ? kb : 0; took = true; }
				}
				if (!took) other_max = q.msc0;
				if (q.plus0 < plus) plus = q.plus0;
			}
		}
		if (simple) {
			const int64_t cL = msc_i - lb + 1;
			if (A.P.mcopy_num > 1 && cL >= A.P.mcopy_khit_cutoff && other_max != INT64_MIN) { // a secondary chain may qualify (Hash_Table.cpp:2184-2190)
				int64_t min_sc = (int64_t)((double)(msc - plus) * A.P.mcopy_rate);
				if (other_max - plus >= min_sc) simple = false;
			}
		}
		if (simple) { // the chain is the whole best block, already contiguous in the grouped buffer: no copy
			const int32_t cL = (int32_t)(msc_i - lb + 1);
			if (lane == 0) {
				hb_chain_t z; hb_hit_t hb = hit_ld(a + lb), he = hit_ld(a + msc_i);
				hb_push_chain(z, xl, yl, msc, hb, he);
				z.first_hit = (uint32_t)(ab + lb); z.n_hits = (uint32_t)cL; z.pad = HB_CHAIN_INPLACE;
				if (A.fc) { FcOut fc; fc.n = 0; fc.ovf = 0; fc.buf = A.fc + ab + 2 * cb; fc.cap = d.count + 2 * ns; hb_gen_fcigar(fc, z, a + lb, cL); if (fc.ovf) atomicOr(A.err, 16); }
				A.ch[cb] = z;
				for (int32_t s = 1; s < ns; s++) A.ch[cb + s].n_hits = 0;
				for (int32_t s = 0; s < ns; s++) A.slot_read[cb + s] = d.read;
			}
		} else {
			// the general case: quick check with f/p written out, warp-parallel DP over the
			// unresolved range, then backtrack / secondary chains / emission on lane 0
			int32_t *f = A.f + ab, *p = A.p + ab, *ii = A.ii + ab; int64_t *t = A.t + ab;
			if (lane == 0 && A.dbg) atomicAdd(&A.dbg[1], 1ull);
			ordered = warp_group_ordered(a, an, lane, &kb);
			if (!ordered) { if (lane == 0) { hb_order_group(a, an); if (A.dbg) atomicAdd(&A.dbg[0], 1ull); } __syncwarp(); warp_group_ordered(a, an, lane, &kb); }
			for (int32_t z = lane; z < an; z += 32) { t[z] = 0; ii[z] = 0; }
			ChainState S; S.plus = 0; S.msc = S.msc_i = INT32_MIN; S.movl = INT32_MAX; S.si = 0; S.ei = an;
			for (int b = 0; b < 2; b++) { // quick_ck_lchain, Hash_Table.cpp:2015-2093
				int32_t l = b ? kb : 0, k = b ? an : kb;
				if (l >= k) continue;
				QBlock q = warp_quick_block(a, l, k, A.P, xl, yl, lane, f, p);
				if (!q.resolved) continue;
				if (q.msc0 >= S.msc) {
					hb_hit_t e = hit_ld(a + q.msc_i0);
					int64_t movl0 = hb_chain_len(e.self_offset, e.self_offset, xl, e.offset, e.offset, yl);
					if (q.msc0 > S.msc || movl0 < S.movl) { S.msc = q.msc0; S.msc_i = q.msc_i0; S.movl = movl0; }
				}
				if (q.plus0 < S.plus) S.plus = q.plus0;
				if (S.ei > k) S.si = k; else S.ei = l;
			}
			__syncwarp();
			warp_chain_dp(a, kb, f, p, t, ii, A.P, xl, yl, S, lane);
			__syncwarp();
			if (lane == 0) {
				FcOut fc; fc.n = 0; fc.ovf = 0; fc.buf = A.fc ? A.fc + ab + 2 * cb : 0; fc.cap = d.count + 2 * ns;
				for (int32_t s = 0; s < ns; s++) A.ch[cb + s].n_hits = 0;
				hb_chain_finish(a, an, A.chits + ab, ab, f, p, t, ii, A.P, xl, yl, A.ch + cb, ns, fc, S);
				for (int32_t s = 0; s < ns; s++) A.slot_read[cb + s] = d.read;
				if (fc.ovf) atomicOr(A.err, 16);
			}
		}
		__syncwarp();
	}
}

// ----------------------------------------------------------------------------
// post: one thread per read
// ----------------------------------------------------------------------------
struct PostArgs {
	DevReads R; uint64_t r0, nR; const uint64_t *c_off; hb_chain_t *ch; const hb_hit_t *chits, *ghits; uint32_t *idx; uint32_t *n_ol; uint8_t *keep;
	uint64_t *cc; const uint64_t *cc_off; ChainPar P;
};

// post, warp per read: the read's chain slots are staged in shared memory (48 B each,
// 128-bit copies by the whole warp), lane 0 runs the sequential post-filter on the staged
// copy (shared-memory latency instead of a dependent L2 access per comparison), the warp
// writes the result back.  Reads with more slots than fit run on global memory as before.
#define POST_WARPS 4
#define POST_SMEM_PER_WARP 16384
__global__ void __launch_bounds__(POST_WARPS * 32) k_post_warp(PostArgs A)
{
	extern __shared__ uint4 post_smem[];
	const int wib = threadIdx.x >> 5, lane = hb_lane();
	const uint64_t r = (uint64_t)blockIdx.x * POST_WARPS + wib;
	if (r >= A.nR) return;
	const uint64_t cb = A.c_off[r]; const uint32_t ns = (uint32_t)(A.c_off[r + 1] - cb);
	uint4 *sw = post_smem + (size_t)wib * (POST_SMEM_PER_WARP / 16);
	hb_chain_t *ch = A.ch + cb; uint32_t *idx = A.idx + cb; uint32_t n = 0;
	if ((size_t)ns * (sizeof(hb_chain_t) + 4) + 16 <= POST_SMEM_PER_WARP) {
		hb_chain_t *s_ch = (hb_chain_t *)sw; uint32_t *s_idx = (uint32_t *)(s_ch + ns);
		const uint4 *src = (const uint4 *)ch;
		for (uint32_t i = lane; i < ns * 3; i += 32) sw[i] = src[i];
		__syncwarp();
		if (lane == 0) n = hb_chain_post(s_ch, ns, A.chits, A.ghits, s_idx, A.cc + A.cc_off[r], A.R.len[A.r0 + r], A.P);
		n = __shfl_sync(HB_FULL, n, 0);
		__syncwarp();
		for (uint32_t i = lane; i < ns; i += 32) ch[i].pad = s_ch[i].pad;
		for (uint32_t i = lane; i < n; i += 32) { const uint32_t s = s_idx[i]; idx[i] = s; A.keep[cb + s] = 1; }
	} else {
		if (lane == 0) n = hb_chain_post(ch, ns, A.chits, A.ghits, idx, A.cc + A.cc_off[r], A.R.len[A.r0 + r], A.P);
		n = __shfl_sync(HB_FULL, n, 0);
		__syncwarp();
		for (uint32_t i = lane; i < n; i += 32) A.keep[cb + idx[i]] = 1;
	}
	if (lane == 0) A.n_ol[r] = n;
}

// ----------------------------------------------------------------------------
// exact: warp per chain slot.  exact_ec_check (ecovlp.cpp:2803) of the query
// interval against the strand-oriented target interval, 16 bases (one 32-bit
// word of 2-bit codes) per lane per step, straight from the packed reads.
// ----------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t ex_fwd16(const uint8_t *p, uint64_t pos)
{ // 16 bases starting at `pos`, first base in the top two bits
	const uint8_t *b = p + (pos >> 2); uint32_t sh = (uint32_t)(pos & 3) << 1;
	uint64_t v = (uint64_t)b[0] << 32 | (uint64_t)b[1] << 24 | (uint64_t)b[2] << 16 | (uint64_t)b[3] << 8 | b[4];
	return (uint32_t)(v >> (8 - sh));
}
static __device__ __forceinline__ uint32_t ex_rc16(uint32_t w)
{ // reverse the 16 2-bit groups and complement
	w = __brev(w); w = ((w >> 1) & 0x55555555u) | ((w & 0x55555555u) << 1);
	return ~w;
}
struct ExactArgs { DevReads R; uint64_t r0, n_slots; const hb_chain_t *ch; const uint32_t *slot_read; const uint8_t *keep; uint8_t *exact; };
__global__ void k_exact(ExactArgs A)
{
	uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = hb_lane();
	if (s >= A.n_slots || !A.keep[s]) return;
	const hb_chain_t c = A.ch[s];
	uint64_t qid = A.r0 + A.slot_read[s], tid = c.y_id, qs = c.x_pos_s, qe = (uint64_t)c.x_pos_e + 1, ts = c.y_pos_s, te = (uint64_t)c.y_pos_e + 1;
	int ok = 1;
	if (qe - qs != te - ts) ok = 0;
	else {
		const uint8_t *q = A.R.packed + A.R.off[qid], *t = A.R.packed + A.R.off[tid];
		uint64_t n = qe - qs, tl = A.R.len[tid], nw = n >> 4;
		for (uint64_t w = lane; w < nw && ok; w += 32) {
			uint32_t a = ex_fwd16(q, qs + (w << 4)), b;
			if (!c.y_pos_strand) b = ex_fwd16(t, ts + (w << 4));
			else b = ex_rc16(ex_fwd16(t, tl - (ts + (w << 4)) - 16));
			if (a != b) ok = 0;
		}
		for (uint64_t i = (nw << 4) + lane; i < n && ok; i += 32) {
			int a = hb_base(q, qs + i), b = c.y_pos_strand ? 3 - hb_base(t, tl - 1 - (ts + i)) : hb_base(t, ts + i);
			if (a != b) ok = 0;
		}
		ok = __all_sync(HB_FULL, ok);
		if (ok && (A.R.noff[qid] != A.R.noff[qid + 1] || A.R.noff[tid] != A.R.noff[tid + 1])) { // N bases: rare, one lane
			if (lane == 0) ok = hb_exact_seq(A.R, qid, qs, qe, tid, ts, te, (int)c.y_pos_strand);
			ok = __shfl_sync(HB_FULL, ok, 0);
		}
	}
	ok = __all_sync(HB_FULL, ok);
	if (lane == 0) A.exact[s] = (uint8_t)ok;
}

// ----------------------------------------------------------------------------
// merge: one thread per read
// ----------------------------------------------------------------------------
struct MergeArgs {
	DevReads R; uint64_t r0, nR; const uint64_t *c_off; const hb_chain_t *ch; const uint32_t *idx; const uint32_t *n_ol; const uint8_t *exact;
	hb_ma_hit_t *in0; const uint64_t *in0_off; const hb_ma_hit_t *in1; const uint64_t *in1_off; // offsets indexed by batch-local read, relative to in0/in1
	FinOv *ov; uint64_t *srt; const uint64_t *o_off; // o_off[r] = sum(slots + n0) before r
	hb_ma_hit_t *out0, *out1; uint32_t *m0, *m1; unsigned long long *stat;
};

// merge, warp per read: the working list (FinOv) and the sort keys are staged in shared
// memory; lane 0 runs the sequential merge there.
#define MERGE_WARPS 4
#define MERGE_SMEM_PER_WARP 14336
#define MERGE_RS_BYTES (2 * 256 * 4 + HB_RS_STACK * 12) /* radix-sort scratch, first in the warp's slice */
__global__ void __launch_bounds__(MERGE_WARPS * 32) k_merge_warp(MergeArgs A)
{
	extern __shared__ uint4 merge_smem[];
	const int wib = threadIdx.x >> 5, lane = hb_lane();
	const uint64_t r = (uint64_t)blockIdx.x * MERGE_WARPS + wib;
	if (r >= A.nR) return;
	if (lane != 0) return;
	const uint64_t cb = A.c_off[r], ob = A.o_off[r], i0 = A.in0_off[r], i1 = A.in1_off[r];
	const uint32_t n0 = (uint32_t)(A.in0_off[r + 1] - i0), n1 = (uint32_t)(A.in1_off[r + 1] - i1), n_ol = A.n_ol[r]; uint32_t m0, m1;
	unsigned long long st[7];
	uint8_t *sw = (uint8_t *)(merge_smem + (size_t)wib * (MERGE_SMEM_PER_WARP / 16));
	RsScratch W = { (int32_t *)sw, (int32_t *)sw + 256, (RsFrame *)(sw + 2048) };
	uint8_t *rest = sw + ((MERGE_RS_BYTES + 15) & ~15);
	FinOv *ov = A.ov + ob; uint64_t *srt = A.srt + i0 + i1;
	if ((size_t)(n_ol + n0) * sizeof(FinOv) + (size_t)(n0 + n1) * 8 + 16 + ((MERGE_RS_BYTES + 15) & ~15) <= MERGE_SMEM_PER_WARP) { srt = (uint64_t *)rest; ov = (FinOv *)(srt + n0 + n1); } // u64 keys first: FinOv is 36 bytes
	hb_final_merge(A.R, (uint32_t)(A.r0 + r), A.ch + cb, A.idx + cb, n_ol, A.exact + cb, A.in0 + i0, n0, A.in1 + i1, n1, ov, srt, A.out0 + ob, &m0, A.out1 + ob, &m1, st, W);
	A.m0[r] = m0; A.m1[r] = m1;
	for (int b = 0; b < 7; b++) if (st[b]) atomicAdd(&A.stat[b], st[b]);
}

// per-read slices (at o_off) -> dense arrays (at d_off); warp per read
__global__ void k_gather_ma(uint64_t nR, const uint64_t *__restrict__ o_off, const uint64_t *__restrict__ d_off, const hb_ma_hit_t *__restrict__ in, hb_ma_hit_t *__restrict__ out)
{
	uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= nR) return;
	const uint4 *s = (const uint4 *)(in + o_off[w]); uint4 *d = (uint4 *)(out + d_off[w]);
	uint32_t n = (uint32_t)(d_off[w + 1] - d_off[w]) * 3; // 48-byte records = 3 x 16 B
	for (uint32_t i = hb_lane(); i < n; i += 32) d[i] = s[i];
}

// ----------------------------------------------------------------------------
// window pass of an EC round: one thread per (chain, window).  The pattern (target
// slice) and the text (query window) are read straight from the 2-bit packed reads;
// the band (<= 63 bits) lives in one 64-bit register pair (row a8).
// ----------------------------------------------------------------------------
struct WinDesc { uint32_t read, slot, k, ord; }; // batch-local read, chain slot, window number, chain ordinal in the read's list
struct WinArgs {
	DevReads R; uint64_t r0, n_win; const WinDesc *desc; const hb_chain_t *ch; const uint64_t *fc; const uint64_t *fc_grp_base;
	double e_rate; int32_t w_l; hb_win_t *out; int *err;
	const uint8_t *ea; const uint64_t *o_off; // row a12 flags per overlap (NULL: none) and the overlap offset of every read: windows of overlaps accepted by the exact shortcut are not aligned
};
static __device__ __forceinline__ uint64_t hb_bswap64(uint64_t v) { return ((uint64_t)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123); }
static __device__ __forceinline__ bool dev_is_n(const DevReads &R, uint64_t rid, uint32_t pos)
{ // membership in the read's (ascending) N list
	uint64_t lo = R.noff[rid], hi = R.noff[rid + 1];
	while (lo < hi) { uint64_t mid = (lo + hi) >> 1; uint32_t v = R.npos[mid]; if (v < pos) lo = mid + 1; else hi = mid; }
	return lo < R.noff[rid + 1] && R.npos[lo] == pos;
}
__global__ void __launch_bounds__(128) k_windows(WinArgs A)
{
	const uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (wi >= A.n_win) return;
	const WinDesc d = A.desc[wi];
	if (A.ea && A.ea[A.o_off[d.read] + d.ord]) return; // step A never reads the windows of an overlap the exact shortcut accepted
	const hb_chain_t c = A.ch[d.slot];
	const uint64_t qid = A.r0 + d.read, tid = c.y_id;
	const uint64_t *fc = A.fc + A.fc_grp_base[d.slot] + c.fc_off;
	const int64_t w_l = A.w_l, xs = c.x_pos_s, xe = c.x_pos_e, n_s = (xs / w_l) * w_l;
	int64_t q_s = n_s + (int64_t)d.k * w_l, q_e = q_s + w_l - 1;
	if (q_s < xs) q_s = xs;
	if (q_e > xe) q_e = xe;
	const int64_t q_l = 1 + q_e - q_s, t_tot_l = A.R.len[tid];
	int64_t thre = (int64_t)((double)q_l * A.e_rate); // Correct.cpp:12962-12964
	if (thre == 0 && q_l >= 4) thre = 1;
	if (thre > 31) thre = 31;
	int64_t t_s = (q_s - xs) + c.y_pos_s, sh = 0;
	{ // y_start_offset, Hash_Table.h:165-189
		const uint32_t n = c.fc_n; bool bad = false;
		if (q_s == (int64_t)(fc[n - 1] >> 32)) sh = hb_fc_shift(fc[n - 1]);
		else {
			uint32_t i = 0;
			for (; i < n; i++) if (q_s < (int64_t)(fc[i] >> 32)) break;
			if (i == 0 || i == n) bad = true; else sh = hb_fc_shift(fc[i - 1]);
		}
		if (bad) atomicOr(A.err, 32);
	}
	t_s += sh;
	hb_win_t rec; rec.chain = (int32_t)d.ord; rec.q_s = (int32_t)q_s; rec.q_e = (int32_t)q_e; rec.t_s = (int32_t)t_s; rec.t_pri_l = -1; rec.thre = (int32_t)thre;
	rec.aux_beg = rec.aux_end = 0; rec.err = INT32_MAX; rec.pe = -1;
	const int64_t aln_l = q_l + (thre << 1);
	// init_waln, Correct.cpp:764-780
	if (!(t_s < 0 || t_s >= t_tot_l || (t_tot_l - t_s + 2 * thre + 31) < aln_l)) {
		int64_t aux_beg = 0, r_s = t_s - thre, r_l = t_tot_l - r_s; if (r_l > aln_l) r_l = aln_l;
		const int64_t aux_end = aln_l - r_l;
		if (r_s < 0) { aux_beg = -r_s; r_s = 0; r_l -= aux_beg; }
		rec.t_s = (int32_t)r_s; rec.t_pri_l = (int32_t)r_l; rec.aux_beg = (int32_t)aux_beg; rec.aux_end = (int32_t)aux_end;
		// ed_band_cal_semi_64_w_absent_diag(pattern = target[r_s, r_s+r_l) on the chain's strand, text = query[q_s, q_e])
		const uint8_t *qp = A.R.packed + A.R.off[qid], *tp = A.R.packed + A.R.off[tid];
		const bool q_has_n = A.R.noff[qid] != A.R.noff[qid + 1], t_has_n = A.R.noff[tid] != A.R.noff[tid + 1], rev = c.y_pos_strand != 0;
		auto pat = [&](int32_t j) -> int { // pattern base j (4 = N)
			const uint32_t fp = rev ? (uint32_t)(t_tot_l - 1 - (r_s + j)) : (uint32_t)(r_s + j);
			if (t_has_n && dev_is_n(A.R, tid, fp)) return 4;
			const int b = hb_base(tp, fp); return rev ? 3 - b : b;
		};
		auto txt = [&](int32_t i) -> int {
			const uint32_t fp = (uint32_t)(q_s + i);
			if (q_has_n && dev_is_n(A.R, qid, fp)) return 4;
			return hb_base(qp, fp);
		};
		const int32_t pn = (int32_t)r_l, tn = (int32_t)q_l, th = (int32_t)thre, abs_diag = (int32_t)aux_beg;
		uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN, X, D0, HN, HP, mm;
		int32_t bd, i, err = abs_diag, i_bd, tn0 = tn - 1, cut = th + (th << 1), best = INT32_MAX, pe = -1, site, ai, uge = INT32_MAX, chh;
		bool dead = pn > tn + cut || tn > pn + cut;
		if (!dead && !q_has_n && !t_has_n) {
			// no N on either read (the rule): the band's match mask comes from two bit planes of the pattern (low / high bit of the base) and a
			// validity plane, shifted with the band — no Peq table, no indexed registers — and both reads are streamed 32 bases per 64-bit load
			const uint64_t *qw = (const uint64_t *)qp, *tw = (const uint64_t *)tp;
			uint64_t tb; int32_t tleft; // text: base q_s + i in the top two bits of tb
			{ const uint64_t p0 = (uint64_t)q_s; tb = hb_bswap64(__ldg(qw + (p0 >> 5))) << (2 * (p0 & 31)); tleft = 32 - (int32_t)(p0 & 31); }
			uint64_t pb; int32_t pleft; int64_t pw;       // pattern: forward -> next base in the top two bits; reverse strand -> next base in the low two bits (complemented)
			if (!rev) { const uint64_t p0 = (uint64_t)r_s; pw = (int64_t)(p0 >> 5); pb = hb_bswap64(__ldg(tw + pw)) << (2 * (p0 & 31)); pleft = 32 - (int32_t)(p0 & 31); }
			else { const uint64_t p0 = (uint64_t)(t_tot_l - 1 - r_s); pw = (int64_t)(p0 >> 5); pb = hb_bswap64(__ldg(tw + pw)) >> (2 * (31 - (p0 & 31))); pleft = (int32_t)(p0 & 31) + 1; }
			auto pnext = [&]() -> uint32_t {
				if (pleft == 0) { if (!rev) { ++pw; pb = hb_bswap64(__ldg(tw + pw)); } else { --pw; pb = hb_bswap64(__ldg(tw + pw)); } pleft = 32; }
				uint32_t b; if (!rev) { b = (uint32_t)(pb >> 62); pb <<= 2; } else { b = 3u - (uint32_t)(pb & 3ULL); pb >>= 2; }
				--pleft; return b;
			};
			uint64_t lo = 0, hi = 0, va = 0;
			bd = ((th << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn;
			for (i = 0; i < bd; i++) { const uint32_t b = pnext(); const int sh = abs_diag + i; lo |= (uint64_t)(b & 1) << sh; hi |= (uint64_t)(b >> 1) << sh; va |= 1ULL << sh; }
			i_bd = (th << 1) - abs_diag; VN = (1ULL << abs_diag) - 1; const int top = th << 1;
			for (i = 0; i <= tn0; i++) {
				if (tleft == 0) { tb = hb_bswap64(__ldg(qw + ((uint64_t)(q_s + i) >> 5))); tleft = 32; }
				const uint32_t tc = (uint32_t)(tb >> 62); tb <<= 2; --tleft;
				const uint64_t Lm = 0ULL - (uint64_t)(tc & 1), Hm = 0ULL - (uint64_t)(tc >> 1);
				X = (va & ~((lo ^ Lm) | (hi ^ Hm))) | VN;
				D0 = ((VP + (X & VP)) ^ VP) | X;
				HN = VP & D0; HP = VN | ~(VP | D0);
				X = D0 >> 1;
				VN = X & HP; VP = HN | ~(X | HP);
				if (!(D0 & 1ULL)) { ++err; if (err > cut) { dead = true; break; } }
				if (i == tn0) break;
				lo >>= 1; hi >>= 1; va >>= 1;
				++i_bd;
				if (i_bd < pn) { const uint32_t b = pnext(); lo |= (uint64_t)(b & 1) << top; hi |= (uint64_t)(b >> 1) << top; va |= 1ULL << top; }
			}
		} else if (!dead) {
			bd = ((th << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn;
			for (i = 0, mm = 1ULL << abs_diag; i < bd; i++) { Peq[pat(i)] |= mm; mm <<= 1; }
			i_bd = (th << 1) - abs_diag; VN = (1ULL << abs_diag) - 1;
			Peq[4] = 0; mm = 1ULL << (th << 1);
			for (i = 0; i <= tn0; i++) {
				X = Peq[txt(i)] | VN;
				D0 = ((VP + (X & VP)) ^ VP) | X;
				HN = VP & D0; HP = VN | ~(VP | D0);
				X = D0 >> 1;
				VN = X & HP; VP = HN | ~(X | HP);
				if (!(D0 & 1ULL)) { ++err; if (err > cut) { dead = true; break; } }
				if (i == tn0) break;
				Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
				++i_bd; chh = 4;
				if (i_bd < pn) chh = pat(i_bd);
				if (chh < 4) Peq[chh] |= mm;
			}
		}
		if (!dead) {
			site = tn - 1 - abs_diag; ai = pn - tn + abs_diag;
			for (i = 0; site < 0 && i < ai; i++, site++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
			if (err <= th && err <= best) { best = err; pe = site; }
			site -= i;
			while (i < ai) {
				err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); ++i;
				if (err <= th && err <= best) { best = err; pe = site + i; }
				if (i == th) uge = err;
			}
			if (uge <= th && uge == best) pe = site + th;
		}
		rec.err = best; rec.pe = pe;
	} else rec.t_s = -1; // init_waln resets its r_s to -1 when it rejects the window (Correct.cpp:766)
	A.out[wi] = rec;
}

// windows per read (thread per read) and their descriptors
__global__ void k_win_count(uint64_t nR, const uint64_t *__restrict__ c_off, const hb_chain_t *__restrict__ ch, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ n_ol, int32_t w_l, uint32_t *__restrict__ wcnt)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	uint64_t cb = c_off[r]; uint32_t tot = 0;
	for (uint32_t i = 0; i < n_ol[r]; i++) { const hb_chain_t &c = ch[cb + idx[cb + i]]; int64_t nl = ((int64_t)c.x_pos_e + 1) - ((int64_t)c.x_pos_s / w_l) * w_l; tot += (uint32_t)(nl / w_l + (nl % w_l > 0 ? 1 : 0)); } // get_num_wins, Correct.cpp:783
	wcnt[r] = tot;
}
__global__ void k_win_desc(uint64_t nR, const uint64_t *__restrict__ c_off, const hb_chain_t *__restrict__ ch, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ n_ol, int32_t w_l, const uint64_t *__restrict__ w_off, WinDesc *__restrict__ desc)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	uint64_t cb = c_off[r], o = w_off[r];
	for (uint32_t i = 0; i < n_ol[r]; i++) {
		const uint32_t s = idx[cb + i]; const hb_chain_t &c = ch[cb + s];
		int64_t nl = ((int64_t)c.x_pos_e + 1) - ((int64_t)c.x_pos_s / w_l) * w_l; uint32_t nw = (uint32_t)(nl / w_l + (nl % w_l > 0 ? 1 : 0));
		for (uint32_t k = 0; k < nw; k++) { WinDesc d; d.read = (uint32_t)r; d.slot = (uint32_t)(cb + s); d.k = k; d.ord = i; desc[o++] = d; }
	}
}

// ----------------------------------------------------------------------------
// step A of the alignment stage of an EC round (rows a8 + a9): one thread per overlap consumes the window records
// of k_windows (hb_ecaln.cuh).  Grid-stride over the overlaps so that the per-thread trace scratch (5 words per
// column of one window) is bounded by the grid, not by the batch.
// ----------------------------------------------------------------------------
// OvDesc (batch-local read, chain slot, number of windows, first window): hb_ecaln.cuh
__global__ void k_ov_desc(uint64_t nR, const uint64_t *__restrict__ c_off, const hb_chain_t *__restrict__ ch, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ n_ol, int32_t w_l,
                          const uint64_t *__restrict__ w_off, const uint64_t *__restrict__ o_off, OvDesc *__restrict__ desc)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	uint64_t cb = c_off[r], w = w_off[r], o = o_off[r];
	for (uint32_t i = 0; i < n_ol[r]; i++) {
		const uint32_t s = idx[cb + i]; const hb_chain_t &c = ch[cb + s];
		int64_t nl = ((int64_t)c.x_pos_e + 1) - ((int64_t)c.x_pos_s / w_l) * w_l; uint32_t nw = (uint32_t)(nl / w_l + (nl % w_l > 0 ? 1 : 0));
		OvDesc d; d.read = (uint32_t)r; d.slot = (uint32_t)(cb + s); d.nw = nw; d.pad = 0; d.w0 = w; desc[o + i] = d; w += nw;
	}
}
#endif // HB_KERNELS_MAIN
struct EcAlnArgs {
	const uint8_t *ea; // row a12 flags (NULL: none)
	DevReads R; uint64_t r0, n_ov; const OvDesc *desc; const hb_chain_t *ch; const uint64_t *fc; const uint64_t *fc_grp_base; const hb_win_t *win;
	double e_rate; int32_t w_l; hb_wl_t *wl; hb_aln_t *out; uint64_t *path; uint16_t *cig_tmp; uint16_t *pool; unsigned long long *pool_used; uint64_t pool_cap; int *err;
	uint32_t *q, *q_n; // overlaps that need the aligner (k_ec_overlap_fast -> k_ec_overlap)
};
#ifdef HB_KERNELS_ECB
// step A in two launches.  k_ec_overlap_fast, thread / overlap at full occupancy: an overlap whose windows ALL aligned in the window pass (the rule) needs no
// aligner — push_hc_wlst_exz finds no gap to fill and gen_extend_err_exz no stretch to estimate — so its window list and error total are sums over the
// records; the others are queued.  k_ec_overlap, thread / queued overlap, grid bounded by the per-thread trace scratch: gap filling and extension estimates.
__global__ void __launch_bounds__(128) k_ec_overlap_fast(EcAlnArgs A)
{
	const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (o >= A.n_ov) return;
	const OvDesc d = A.desc[o];
	hb_aln_t res; res.w_off = d.w0; res.pad = 0;
	if (A.ea && A.ea[o]) { res.st = 3; res.align_length = 0; res.rr = 0; res.re = 0; res.w_n = 0; A.out[o] = res; return; } // accepted by the previous round's exact record: no alignment
	bool all = true;
	for (uint32_t k = 0; k < d.nw; k++) { const hb_win_t &wr = A.win[d.w0 + k]; if (wr.t_pri_l < 0 || wr.err > wr.thre) { all = false; break; } }
	if (!all) { A.q[atomicAdd(A.q_n, 1u)] = (uint32_t)o; return; }
	const hb_chain_t c = A.ch[d.slot];
	EcCtx C; C.R = A.R; C.e_rate = A.e_rate; C.w_l = A.w_l; C.pool = A.pool; C.pool_used = A.pool_used; C.pool_cap = A.pool_cap; C.err = A.err;
	C.ez.path = 0; C.ez.cig = 0; C.ez.cn = 0;
	C.q = hb_rd_view(A.R, A.r0 + d.read, 0); C.t = hb_rd_view(A.R, c.y_id, c.y_pos_strand);
	hb_ec_overlap_A(C, c, A.fc + A.fc_grp_base[d.slot] + c.fc_off, A.win + d.w0, (int32_t)d.nw, A.wl + d.w0, &res);
	A.out[o] = res;
}
__global__ void __launch_bounds__(64) k_ec_overlap(EcAlnArgs A)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
	EcCtx C; C.R = A.R; C.e_rate = A.e_rate; C.w_l = A.w_l; C.pool = A.pool; C.pool_used = A.pool_used; C.pool_cap = A.pool_cap; C.err = A.err;
	C.ez.path = A.path + tid * (uint64_t)A.w_l * 5; C.ez.cig = A.cig_tmp + tid * HB_EC_CIG_TMP; C.ez.cn = 0;
	const uint64_t n_work = *A.q_n;
	for (uint64_t wk = tid; wk < n_work; wk += nthr) {
		const uint64_t o = A.q[wk];
		const OvDesc d = A.desc[o]; const hb_chain_t c = A.ch[d.slot];
		C.q = hb_rd_view(A.R, A.r0 + d.read, 0); C.t = hb_rd_view(A.R, c.y_id, c.y_pos_strand);
		hb_aln_t res; res.w_off = d.w0; res.pad = 0;
		hb_ec_overlap_A(C, c, A.fc + A.fc_grp_base[d.slot] + c.fc_off, A.win + d.w0, (int32_t)d.nw, A.wl + d.w0, &res);
		A.out[o] = res;
	}
}

#endif // HB_KERNELS_ECB
// ----------------------------------------------------------------------------
// steps B + C (rows a10 + a11): base-level CIGAR of the accepted overlaps as three kernels (hb_ecaln.cuh):
//   k_ecb_prep   thread / overlap : chain refinement, whole-overlap exact shortcut, segment count
//   k_ecb_seg    thread / segment : the alignments (independent of each other); scratch tiers chained by queues
//   k_ecb_merge  thread / overlap : window fusion over the stored results, indel normalisation, totals
// ----------------------------------------------------------------------------
#ifdef HB_KERNELS_MAIN
__global__ void k_ecb_cap(uint64_t n_ov, const OvDesc *__restrict__ desc, const hb_chain_t *__restrict__ ch, const hb_aln_t *__restrict__ aln, uint32_t *__restrict__ cap)
{ // window capacity of an overlap = number of inter-anchor segments (+1 spare)
	uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (o >= n_ov) return;
	cap[o] = aln[o].st == 2 ? ch[desc[o].slot].n_hits + 2 + HB_RC_SPARE_WIN : (aln[o].st == 3 ? 2 : 0); // the spare: windows the re-seeding rescue may add
}
#endif // HB_KERNELS_MAIN
struct EcCigArgs {
	DevReads R; uint64_t r0, n_ov; const OvDesc *desc; const hb_chain_t *ch; const uint64_t *fc; const uint64_t *fc_grp_base;
	const hb_aln_t *aln; const hb_wl_t *wlA; hb_hit_t *chits, *ghits; uint64_t dp_half; int64_t *dp_t, *dp_p; int32_t *dp_f;
	double e_rate; int32_t w_l; int gaps;
	EcPrep *prep; uint32_t *nseg; const uint64_t *seg_off; uint64_t n_seg; EcSeg *segs; // segments of overlap o: segs[seg_off[o] .. seg_off[o+1])
	uint16_t *spool; unsigned long long *spool_used; uint64_t spool_cap;                // cigars of the aligned segments
	uint32_t *q_in; const uint32_t *q_in_n; uint32_t *q_out; uint32_t *q_out_n; uint32_t *work, *q_key; // deferred segments (ids) between scratch tiers; work: the warp tiers' dynamic counter
	uint64_t *path; uint64_t path_words; uint64_t *vec; int32_t vstride; uint16_t *cig_tmp; int32_t cig_words; // per-thread scratch of the queue tiers / merge
	int pass; hb_alnb_t *out; hb_wl_t *wl; const uint64_t *wl_off;
	uint16_t *pool; unsigned long long *pool_used; uint64_t pool_cap; unsigned int *n_deferred; int *err;
	uint32_t *rc_q; unsigned int *rc_n; // overlaps that ask for the re-seeding rescue (k_ecb_rechain, ecrechain.cu)
};
#ifdef HB_KERNELS_ECB
static __device__ __forceinline__ void ecb_overlap_view(const EcCigArgs &A, uint64_t o, const hb_aln_t &a, OvDesc &d, hb_chain_t &c, EcZ &z, hb_hit_t *&ch_a, uint64_t &dpo)
{
	d = A.desc[o]; c = A.ch[d.slot];
	const bool inpl = (c.pad & HB_CHAIN_INPLACE) != 0;
	ch_a = (inpl ? A.ghits : A.chits) + c.first_hit; dpo = (inpl ? A.dp_half : 0) + c.first_hit;
	z.x_pos_s = c.x_pos_s; z.x_pos_e = c.x_pos_e; z.y_pos_s = c.y_pos_s; z.y_id = c.y_id; z.rev = c.y_pos_strand;
	z.fc = A.fc + A.fc_grp_base[d.slot] + c.fc_off; z.fc_n = c.fc_n; z.align_length = a.align_length; z.w = (hb_wl_t *)(A.wlA + a.w_off); z.wn = (int32_t)a.w_n;
}
// prep: thread / overlap — refine the chain in place, whole-overlap exact shortcut, number of segments
__global__ void __launch_bounds__(128) k_ecb_prep(EcCigArgs A)
{
	const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (o >= A.n_ov) return;
	const hb_aln_t a = A.aln[o]; EcPrep pr; pr.ch_n = 0; pr.shortcut = 0; pr.q0 = pr.q1 = pr.t0 = pr.t1 = 0; uint32_t ns = 0;
	if (a.st == 3) { const hb_chain_t c = A.ch[A.desc[o].slot]; pr.shortcut = 1; pr.q0 = (int32_t)c.x_pos_s; pr.q1 = (int32_t)c.x_pos_e + 1; pr.t0 = (int32_t)c.y_pos_s; pr.t1 = (int32_t)c.y_pos_e + 1; }
	else if (a.st == 2) {
		OvDesc d; hb_chain_t c; EcZ z; hb_hit_t *ch_a; uint64_t dpo;
		ecb_overlap_view(A, o, a, d, c, z, ch_a, dpo);
		hb_ecb_prep(z, a.re, A.R.len[A.r0 + d.read], A.R.len[c.y_id], ch_a, c.n_hits, A.dp_t + dpo, A.dp_p + dpo, A.dp_f + dpo, &pr);
		ns = (pr.shortcut || pr.ch_n <= 0) ? 0 : (uint32_t)pr.ch_n + 1;
	}
	A.prep[o] = pr; A.nseg[o] = ns;
}
// segment pre-pass: thread / inter-anchor segment of the batch.  Everything that needs no alignment is decided here — empty segments,
// segments inside exact windows, segments whose two substrings are identical (16 bases per compare) — and stored; the rest (an error
// inside the segment: ~13 % at 0.2 % read error) is queued for the alignment kernel.  Light on registers, no private arrays.
__global__ void __launch_bounds__(128) k_ecb_seg_fast(EcCigArgs A)
{
	const uint64_t sidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	__shared__ uint64_t s_o;
	if (threadIdx.x == 0) { // overlap of the block's first segment: last o with seg_off[o] <= sidx
		const uint64_t s0 = (uint64_t)blockIdx.x * blockDim.x; uint64_t lo = 0, hi = A.n_ov;
		while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (A.seg_off[mid] <= s0) lo = mid; else hi = mid; }
		s_o = lo;
	}
	__syncthreads();
	bool need = false; uint32_t qlen = 0;
	if (sidx < A.n_seg) {
		uint64_t o = s_o; while (A.seg_off[o + 1] <= sidx) o++;
		const hb_aln_t a = A.aln[o]; const EcPrep pr = A.prep[o];
		OvDesc d; hb_chain_t c; EcZ z; hb_hit_t *ch_a; uint64_t dpo;
		ecb_overlap_view(A, o, a, d, c, z, ch_a, dpo);
		uint16_t one[2];
		EcBCtx C; C.gout = 0; C.gcap = 0; C.e_rate = A.e_rate; C.w_l = A.w_l; C.do_gaps = 0; C.no_myers = 1; C.pool = 0; C.pool_used = 0; C.pool_cap = 0; C.aw = 0; C.awcap = 0; C.wc = 0; C.wccap = 0;
		C.ez.path = 0; C.ez.pcap = 0; C.ez.vec = 0; C.ez.vstride = 0; C.ez.warp = 0; C.ez.cig = one; C.ez.ccap = 2; C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0;
		C.q = hb_rd_view(A.R, A.r0 + d.read, 0); C.t = hb_rd_view(A.R, c.y_id, c.y_pos_strand); C.ql = C.q.len; C.tl = C.t.len;
		int64_t uq[2], ut[2], um;
		const int st = hb_ecb_segment_t<false, true>(C, z, ch_a, pr.ch_n, (int64_t)(sidx - A.seg_off[o]), uq, ut, &um, true);
		if (C.bad) atomicOr(A.err, 32);
		need = st == 5 || C.ez.ovf; // an alignment (or a > 32 k-base exact run) is needed
		{ const int64_t l = uq[1] - uq[0]; qlen = (uint32_t)(l < 0 ? 0 : l > 4095 ? 4095 : l); } // sort key of the queue: lanes of a warp of the alignment kernel then step about the same number of columns
		if (!need) { EcSeg sg; hb_seg_store(C, st, uq, ut, um, &sg, A.spool, A.spool_used, A.spool_cap); A.segs[sidx] = sg; }
	}
	// one queue reservation per warp: the warp's segments stay together and in order, so the alignment kernel's warps mostly work
	// on neighbouring segments of the same overlap (same two reads)
	const unsigned m = __ballot_sync(0xffffffffu, need); const int lane = threadIdx.x & 31;
	if (m) {
		uint32_t base = 0;
		if (lane == __ffs(m) - 1) base = atomicAdd(A.q_out_n, (uint32_t)__popc(m));
		base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
		if (need) { const uint32_t qi = base + __popc(m & ((1u << lane) - 1)); A.q_out[qi] = (uint32_t)sidx; if (A.q_key) A.q_key[qi] = qlen; }
	}
}
// segment alignment, tier 0: one thread per queued segment with a small private scratch (trace of 640 words = 212 columns of a one-word band
// (3 words per column), 4-word band at most: the usual segment between neighbouring minimizers).  A segment that overflows it queues for the warp tiers.
#define ECB_T0_PATH 640
#define ECB_T0_VS 4
#define ECB_T0_CIG 72
__global__ void __launch_bounds__(128) k_ecb_seg(EcCigArgs A)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
	uint64_t l_path[ECB_T0_PATH], l_vec[11 * ECB_T0_VS]; uint16_t l_cig[ECB_T0_CIG];
	EcBCtx C; C.gout = 0; C.gcap = 0; C.e_rate = A.e_rate; C.w_l = A.w_l; C.do_gaps = 0; C.no_myers = 0; C.pool = 0; C.pool_used = 0; C.pool_cap = 0; C.aw = 0; C.awcap = 0; C.wc = 0; C.wccap = 0;
	C.ez.path = l_path; C.ez.pcap = ECB_T0_PATH; C.ez.vec = l_vec; C.ez.vstride = ECB_T0_VS; C.ez.warp = 0; C.ez.cig = l_cig; C.ez.ccap = ECB_T0_CIG;
	const uint64_t n_work = (uint64_t)*A.q_in_n;
	for (uint64_t wk0 = tid - (threadIdx.x & 31); wk0 < n_work; wk0 += nthr) { // warp-uniform trip count: the lanes of a warp stay together (hb_seg_align_t<true> syncs them)
		const uint64_t wk = wk0 + (threadIdx.x & 31); const bool live = wk < n_work;
		uint64_t sidx = 0, o = 0; EcPrep pr; pr.ch_n = 0; OvDesc d; hb_chain_t c; EcZ z; hb_hit_t *ch_a = 0; uint64_t dpo;
		if (live) {
			sidx = A.q_in[wk];
			uint64_t lo = 0, hi = A.n_ov; // overlap of the segment: last o with seg_off[o] <= sidx
			while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (A.seg_off[mid] <= sidx) lo = mid; else hi = mid; }
			o = lo; const hb_aln_t a = A.aln[o]; pr = A.prep[o];
			ecb_overlap_view(A, o, a, d, c, z, ch_a, dpo);
			C.q = hb_rd_view(A.R, A.r0 + d.read, 0); C.t = hb_rd_view(A.R, c.y_id, c.y_pos_strand); C.ql = C.q.len; C.tl = C.t.len;
		}
		C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0;
		int64_t uq[2], ut[2], um;
		const int st = hb_ecb_segment_t<true>(C, z, ch_a, pr.ch_n, (int64_t)(sidx - (live ? A.seg_off[o] : 0)), uq, ut, &um, live);
		if (live) {
			if (C.bad) atomicOr(A.err, 32);
			EcSeg sg; hb_seg_store(C, st, uq, ut, um, &sg, A.spool, A.spool_used, A.spool_cap);
			if (sg.status == 4 && A.q_out) { const uint32_t qi = atomicAdd(A.q_out_n, 1u); A.q_out[qi] = (uint32_t)sidx; }
			A.segs[sidx] = sg;
		}
	}
}
// segment alignment, tier 1: one thread per queued segment, grid-stride, with a launch-sized slice of global scratch (4096 trace words, 8-word band): the
// one- and few-word bands that only outgrew tier 0's private trace (long error-free stretches between distant anchors) — a warp per segment would idle 31 lanes on them
__global__ void __launch_bounds__(128) k_ecb_seg_g(EcCigArgs A)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
	EcBCtx C; C.gout = 0; C.gcap = 0; C.e_rate = A.e_rate; C.w_l = A.w_l; C.do_gaps = 0; C.no_myers = 0; C.pool = 0; C.pool_used = 0; C.pool_cap = 0; C.aw = 0; C.awcap = 0; C.wc = 0; C.wccap = 0;
	C.ez.path = A.path + tid * A.path_words; C.ez.pcap = A.path_words; C.ez.vec = A.vec + tid * 11 * (uint64_t)A.vstride; C.ez.vstride = A.vstride; C.ez.warp = 0; C.ez.cig = A.cig_tmp + tid * (uint64_t)A.cig_words; C.ez.ccap = A.cig_words;
	const uint64_t n_work = (uint64_t)*A.q_in_n;
	for (uint64_t wk0 = tid - (threadIdx.x & 31); wk0 < n_work; wk0 += nthr) { // warp-uniform trip count (see k_ecb_seg)
		const uint64_t wk = wk0 + (threadIdx.x & 31); const bool live = wk < n_work;
		uint64_t sidx = 0, o = 0; EcPrep pr; pr.ch_n = 0; OvDesc d; hb_chain_t c; EcZ z; hb_hit_t *ch_a = 0; uint64_t dpo;
		if (live) {
			sidx = A.q_in[wk];
			uint64_t lo = 0, hi = A.n_ov;
			while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (A.seg_off[mid] <= sidx) lo = mid; else hi = mid; }
			o = lo; const hb_aln_t a = A.aln[o]; pr = A.prep[o];
			ecb_overlap_view(A, o, a, d, c, z, ch_a, dpo);
			C.q = hb_rd_view(A.R, A.r0 + d.read, 0); C.t = hb_rd_view(A.R, c.y_id, c.y_pos_strand); C.ql = C.q.len; C.tl = C.t.len;
		}
		C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0;
		int64_t uq[2], ut[2], um;
		const int st = hb_ecb_segment_t<true>(C, z, ch_a, pr.ch_n, (int64_t)(sidx - (live ? A.seg_off[o] : 0)), uq, ut, &um, live);
		if (live) {
			if (C.bad) atomicOr(A.err, 32);
			EcSeg sg; hb_seg_store(C, st, uq, ut, um, &sg, A.spool, A.spool_used, A.spool_cap);
			if (sg.status == 4 && A.q_out) { const uint32_t qi = atomicAdd(A.q_out_n, 1u); A.q_out[qi] = (uint32_t)sidx; }
			A.segs[sidx] = sg;
		}
	}
}
// segment alignment, warp tiers: one WARP per queued segment, handed out dynamically (A.work).  Every lane runs the reference's scalar logic of the
// segment (coordinates, error estimate, threshold escalation) identically; inside the aligner the band's words are spread over the lanes
// (hb_mwalign_w.cuh) and the trace goes to the warp's slice of a launch-sized scratch.  A segment that overflows the slice queues for the next tier.
__global__ void __launch_bounds__(128) k_ecb_seg_w(EcCigArgs A)
{
	const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
	EcBCtx C; C.gout = 0; C.gcap = 0; C.e_rate = A.e_rate; C.w_l = A.w_l; C.do_gaps = 0; C.no_myers = 0; C.pool = 0; C.pool_used = 0; C.pool_cap = 0; C.aw = 0; C.awcap = 0; C.wc = 0; C.wccap = 0;
	C.ez.path = A.path + wid * A.path_words; C.ez.pcap = A.path_words; C.ez.vec = A.vec + wid * 2 * (uint64_t)A.vstride; C.ez.vstride = A.vstride; C.ez.warp = 1; C.ez.cig = A.cig_tmp + wid * (uint64_t)A.cig_words; C.ez.ccap = A.cig_words;
	const uint32_t n_work = *A.q_in_n;
	for (;;) {
		uint32_t wk = 0; if (lane == 0) wk = atomicAdd(A.work, 1u);
		wk = __shfl_sync(0xffffffffu, wk, 0);
		if (wk >= n_work) break;
		const uint64_t sidx = A.q_in[wk];
		uint64_t lo = 0, hi = A.n_ov;
		while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (A.seg_off[mid] <= sidx) lo = mid; else hi = mid; }
		const uint64_t o = lo; const hb_aln_t a = A.aln[o]; const EcPrep pr = A.prep[o];
		OvDesc d; hb_chain_t c; EcZ z; hb_hit_t *ch_a; uint64_t dpo;
		ecb_overlap_view(A, o, a, d, c, z, ch_a, dpo);
		C.q = hb_rd_view(A.R, A.r0 + d.read, 0); C.t = hb_rd_view(A.R, c.y_id, c.y_pos_strand); C.ql = C.q.len; C.tl = C.t.len;
		C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0;
		int64_t uq[2], ut[2], um;
		const int st = hb_ecb_segment(C, z, ch_a, pr.ch_n, (int64_t)(sidx - A.seg_off[o]), uq, ut, &um);
		__syncwarp();
		if (lane == 0) {
			if (C.bad) atomicOr(A.err, 32);
			EcSeg sg; hb_seg_store(C, st, uq, ut, um, &sg, A.spool, A.spool_used, A.spool_cap);
			if (sg.status == 4 && A.q_out) { const uint32_t qi = atomicAdd(A.q_out_n, 1u); A.q_out[qi] = (uint32_t)sidx; }
			A.segs[sidx] = sg;
		}
		__syncwarp(); // the cigar scratch is read by lane 0 above and rewritten by the next segment
	}
}
// merge: thread / overlap, grid-stride — push_alnw / push_unmap_alnw over the stored segments, reassign_gaps when a window closes,
// totals and update_overlap_region.  pass = 1: only the overlaps the first launch deferred (cigar buffers too small).
__global__ void __launch_bounds__(64) k_ecb_merge(EcCigArgs A)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
	EcBCtx C; C.gout = 0; C.gcap = 0; C.e_rate = A.e_rate; C.w_l = A.w_l; C.pool = A.pool; C.pool_used = A.pool_used; C.pool_cap = A.pool_cap; C.do_gaps = A.gaps;
	uint16_t *base = A.cig_tmp + tid * 3 * (uint64_t)A.cig_words; // wc | gap output | adjust_gap scratch
	C.wc = base; C.wccap = A.cig_words; C.ez.cig = base + A.cig_words; C.ez.ccap = A.cig_words; C.ez.path = (uint64_t *)(base + 2 * (uint64_t)A.cig_words); C.ez.pcap = (uint64_t)A.cig_words / 4;
	C.ez.vec = 0; C.ez.vstride = 0; C.ez.warp = 0;
	for (uint64_t o = tid; o < A.n_ov; o += nthr) {
		const hb_aln_t a = A.aln[o];
		if (A.pass == 0) { if (a.st != 2 && a.st != 3) { hb_alnb_t r; r.st = a.st; r.need_rechain = 0; r.re = 0; r.nh_err = 0; r.x_pos_s = r.x_pos_e = r.y_pos_s = r.y_pos_e = 0; r.w_off = A.wl_off[o]; r.w_n = 0; r.pad = 0; A.out[o] = r; continue; } }
		else if (A.out[o].st != -1) continue;
		OvDesc d; hb_chain_t c; EcZ z; hb_hit_t *ch_a; uint64_t dpo;
		ecb_overlap_view(A, o, a, d, c, z, ch_a, dpo);
		C.q = hb_rd_view(A.R, A.r0 + d.read, 0); C.t = hb_rd_view(A.R, c.y_id, c.y_pos_strand); C.ql = C.q.len; C.tl = C.t.len;
		C.aw = A.wl + A.wl_off[o]; C.awcap = (int32_t)(A.wl_off[o + 1] - A.wl_off[o]);
		hb_alnb_t r; r.w_off = A.wl_off[o]; r.pad = 0;
		hb_ecb_merge(C, z, a.re, A.prep[o], A.segs + A.seg_off[o], A.spool, &r);
		if (a.st == 3 && r.st == 2) { r.x_pos_s = c.x_pos_s; r.x_pos_e = c.x_pos_e; r.y_pos_s = c.y_pos_s; r.y_pos_e = c.y_pos_e; r.pad = 1; } // no update_overlap_region on this path (ecovlp.cpp:2849-2852)
		if (r.st == -1) atomicAdd(A.n_deferred, 1u);
		if (r.st == -2) atomicOr(A.err, 32);
		if (r.st == 2 && r.need_rechain && A.rc_q) A.rc_q[atomicAdd(A.rc_n, 1u)] = (uint32_t)o;
		A.out[o] = r;
	}
}
#endif // HB_KERNELS_ECB
#ifdef HB_KERNELS_MAIN
__global__ void k_gather_wl(uint64_t n_ov, const hb_alnb_t *__restrict__ in, const uint64_t *__restrict__ dense_off, const hb_wl_t *__restrict__ wl, hb_alnb_t *__restrict__ out, hb_wl_t *__restrict__ dense)
{ // capacity-strided window lists -> dense, one thread per overlap (lists are 1-2 entries long in the common case)
	uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (o >= n_ov) return;
	hb_alnb_t r = in[o]; const uint64_t d = dense_off[o];
	for (uint32_t k = 0; k < r.w_n; k++) dense[d + k] = wl[r.w_off + k];
	r.w_off = d; out[o] = r;
}
__global__ void k_ecb_wn(uint64_t n_ov, const hb_alnb_t *__restrict__ in, uint32_t *__restrict__ wn)
{ uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (o < n_ov) wn[o] = in[o].w_n; }

#endif // HB_KERNELS_MAIN
// ----------------------------------------------------------------------------
// phasing (row a13, hb_ecphase.cuh): thread / read.  k_ph_count gathers the read's accepted overlaps, counts mismatch columns
// per query position and sizes the evidence; k_ph_decide collects the evidence, derives the allele statistics and makes the
// greedy haplotype call (is_match 1 / 2, strong) with exactly-sized scratch.
// ----------------------------------------------------------------------------
struct PhArgs {
	DevReads R; uint64_t r0, nR; const uint64_t *o_off; const OvDesc *desc; const hb_chain_t *ch; const hb_aln_t *aln; const hb_alnb_t *alnb; const hb_wl_t *wl; const uint16_t *pool;
	PhOv *ov; uint32_t *n_acc; const uint64_t *b_off; uint8_t *cnt; uint32_t *n_site, *n_ev; const uint64_t *site_off, *ev_off;
	uint32_t *site_pos, *site_o; PhEv *ev, *ev2; PhSnp *snp; uint64_t *ord; uint32_t *ov_o; hb_phase_t *out; int *err; uint32_t *work; // work[2]: dynamic read counters of the two kernels
};
#define HB_CNS_RS_WORDS (512 + 3 * HB_RS_STACK) // dedup_chains' / the phasing's radix-sort scratch (RsScratch) as int32 words
#define PH_WARPS 4
#ifdef HB_KERNELS_ECP
// warp / read, reads handed out dynamically (A.work[0] / A.work[1]): lanes share the cigar walks of the read's alignment windows (hb_ph_count_w / hb_ph_decide_w)
__global__ void __launch_bounds__(PH_WARPS * 32) k_ph_count(PhArgs A)
{
	const int lane = threadIdx.x & 31;
	for (;;) {
		uint32_t r = 0; if (lane == 0) r = atomicAdd(A.work, 1u);
		r = __shfl_sync(0xffffffffu, r, 0);
		if (r >= A.nR) break;
		const uint64_t o0 = A.o_off[r], o1 = A.o_off[r + 1]; PhOv *ov = A.ov + o0; uint32_t n = 0;
		if (lane == 0) for (uint64_t o = o0; o < o1; o++) {
			const hb_alnb_t b = A.alnb[o]; if (b.st != 2) continue;
			const hb_chain_t &c = A.ch[A.desc[o].slot];
			PhOv v; v.w = A.wl + b.w_off; v.wn = b.w_n; v.pool = A.pool; v.y_id = c.y_id; v.rev = c.y_pos_strand; v.align_length = A.aln[o].align_length; v.is_match = 1; v.strong = 0;
			ov[n++] = v;
		}
		n = __shfl_sync(0xffffffffu, n, 0); __syncwarp();
		uint32_t ns = 0, ne = 0;
		if (n) hb_ph_count_w(ov, n, A.cnt + A.b_off[r], A.R.len[A.r0 + r], &ns, &ne);
		if (lane == 0) { A.n_acc[r] = n; A.n_site[r] = ns; A.n_ev[r] = ne; }
	}
}
__global__ void __launch_bounds__(PH_WARPS * 32) k_ph_decide(PhArgs A)
{
	__shared__ int32_t s_rs[PH_WARPS][HB_CNS_RS_WORDS];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	RsScratch W; W.bb = s_rs[warp]; W.be = s_rs[warp] + 256; W.st = (RsFrame *)(s_rs[warp] + 512);
	for (;;) {
		uint32_t r = 0; if (lane == 0) r = atomicAdd(A.work + 1, 1u);
		r = __shfl_sync(0xffffffffu, r, 0);
		if (r >= A.nR) break;
		const uint64_t o0 = A.o_off[r], o1 = A.o_off[r + 1]; PhOv *ov = A.ov + o0; const uint32_t n = A.n_acc[r]; int ovf = 0;
		if (n) hb_ph_decide_w(A.R, A.r0 + r, ov, n, A.cnt + A.b_off[r], A.R.len[A.r0 + r], A.n_site[r], A.n_ev[r], A.site_pos + A.site_off[r], A.site_o + A.site_off[r] + r,
		                      A.ev + A.ev_off[r], A.ev2 + A.ev_off[r], A.snp + 4 * A.site_off[r], A.ord + o0, A.ov_o + o0 + r, W, 3, 3, 0.04, &ovf); // s_hap_cov = infor_cov = 3 (CommandLines.cpp:333-334), up = 0.04
		if (ovf) atomicOr(A.err, 128);
		__syncwarp(); // lane 0's calls are in ov[]
		uint32_t k0 = 0; // records of the read's overlaps: lanes take overlaps, the rank among the accepted ones by a running ballot count
		for (uint64_t ob = o0; ob < o1; ob += 32) {
			const uint64_t o = ob + lane; const bool in = o < o1; hb_alnb_t b; b.st = 0; if (in) b = A.alnb[o];
			const unsigned acc = __ballot_sync(0xffffffffu, in && b.st == 2);
			if (in) {
				const hb_chain_t &c = A.ch[A.desc[o].slot]; hb_phase_t p;
				p.st = b.st; p.y_id = c.y_id; p.rev = c.y_pos_strand; p.is_match = 0; p.strong = 0; p.need_rechain = 0;
				if (b.st == 2) { const uint32_t k = k0 + __popc(acc & ((1u << lane) - 1)); p.x_pos_s = b.x_pos_s; p.x_pos_e = b.x_pos_e; p.y_pos_s = b.y_pos_s; p.y_pos_e = b.y_pos_e; p.nh_err = (uint32_t)b.nh_err; p.is_match = ov[k].is_match; p.strong = ov[k].strong; p.need_rechain = (uint32_t)b.need_rechain; }
				else { p.x_pos_s = c.x_pos_s; p.x_pos_e = c.x_pos_e; p.y_pos_s = c.y_pos_s; p.y_pos_e = c.y_pos_e; p.nh_err = 0; }
				A.out[o] = p;
			}
			k0 += __popc(acc);
		}
		__syncwarp();
	}
}

#endif // HB_KERNELS_ECP
#ifdef HB_KERNELS_MAIN
// the round's reverse_paf[i]: dedup_chains + push_ne_ovlp(flag 2) per read (hb_ecphase.cuh: hb_ec_reverse_list)
__global__ void __launch_bounds__(64) k_ec_rpaf(DevReads R, uint64_t r0, uint64_t nR, const uint64_t *__restrict__ o_off, const hb_phase_t *__restrict__ ph, uint64_t *ord, hb_ma_hit_t *out, uint32_t *n_out, int *err)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	int32_t l_bb[256], l_be[256]; RsFrame l_st[HB_RS_STACK]; RsScratch W = { l_bb, l_be, l_st }; int ovf = 0;
	const uint64_t o0 = o_off[r];
	n_out[r] = hb_ec_reverse_list(R, r0 + r, ph + o0, (uint32_t)(o_off[r + 1] - o0), (PhPair *)(ord + o0), W, out + o0, &ovf);
	if (ovf) atomicOr(err, 128);
}

// the round's paf[i] (row a15): dedup_chains again on its own order array, push_ne_ovlp(flag 1, ec) with extract_max_exact through the read's
// edit script, the large-indel flag, check_well_cal (hb_ecround.cuh).  srt = 2 words per overlap of scratch for the coverage sweep.
__global__ void __launch_bounds__(64) k_ec_spaf(DevReads R, uint64_t r0, uint64_t nR, const uint64_t *__restrict__ o_off, const hb_phase_t *__restrict__ ph, const hb_alnb_t *__restrict__ alnb,
                                                const hb_wl_t *__restrict__ wl, const uint16_t *__restrict__ pool, const uint16_t *__restrict__ sc, const uint64_t *__restrict__ sc_off, uint64_t sc_rid0,
                                                uint64_t *ord, uint64_t *srt, hb_ma_hit_t *out, uint32_t *n_out, uint8_t *flags, int *err)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	int32_t l_bb[256], l_be[256]; RsFrame l_st[HB_RS_STACK]; RsScratch W = { l_bb, l_be, l_st }; int ovf = 0;
	const uint64_t o0 = o_off[r], rid = r0 + r; const uint32_t n = (uint32_t)(o_off[r + 1] - o0);
	const uint16_t *ec = sc + sc_off[rid - sc_rid0]; const uint32_t ecn = (uint32_t)(sc_off[rid - sc_rid0 + 1] - sc_off[rid - sc_rid0]); // sc_off[0] belongs to read sc_rid0
	const uint32_t keep = hb_ec_dedup(ph + o0, n, (PhPair *)(ord + o0), W, &ovf);
	if (ovf) { atomicOr(err, 128); n_out[r] = 0; flags[2 * r] = flags[2 * r + 1] = 0; return; }
	const uint32_t no = hb_ec_source_list(R, rid, ph + o0, alnb + o0, (const PhPair *)(ord + o0), keep, wl, pool, ec, (int64_t)ecn, out + o0);
	n_out[r] = no;
	hb_check_well_cal(ec, ecn, srt + 2 * o0, out + o0, no, R.len[rid], 6 /* MIN_COVERAGE_THRESHOLD * 2, ecovlp.cpp:3324 */, &flags[2 * r], &flags[2 * r + 1]);
}

// window consensus of a batch of reads (row a14, hb_eccns.cuh): one thread per read, grid-stride (the per-thread vote array, 2 * HB_CNS_WL words, is
// bounded by the grid).  ent_cap[r] = windows of the read's accepted overlaps (upper bound of its entries); scripts go to fixed per-read slots.
__global__ void k_cns_cap(uint64_t nR, const uint64_t *__restrict__ o_off, const hb_alnb_t *__restrict__ alnb, uint32_t *__restrict__ ent_cap)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	uint32_t c = 0;
	for (uint64_t j = o_off[r]; j < o_off[r + 1]; j++) if (alnb[j].st == 2) c += alnb[j].w_n;
	ent_cap[r] = c;
}
#endif // HB_KERNELS_MAIN
struct CnsArgs {
	DevReads R; uint64_t r0, nR; const uint64_t *o_off; const hb_phase_t *ph; const hb_alnb_t *alnb; const hb_wl_t *wl; const uint16_t *pool;
	uint64_t *ord; CnsOv *cov; const uint64_t *ent_off; CnsEnt *ent; uint32_t *srt, *act_a, *act_b, *b32; uint64_t *key; int32_t *rs; uint32_t *work; // rs: HB_CNS_RS_WORDS per warp (dedup_chains' sort scratch)
	const uint64_t *out_off; uint16_t *out; uint32_t *out_n; uint8_t *status; unsigned long long *nec; int *err;
	// second launch: the reads the first one reported (queue), each thread with the arena of the graph consensus
	const uint32_t *queue; uint32_t n_queue, g_nodes, g_arcs, g_nseq, g_pcap, g_ccap;
	CnsNode *g_nd; CnsArc *g_arc; uint32_t *g_q, *g_b32, *g_np; uint8_t *g_ns; uint64_t *g_path, *g_vec; uint16_t *g_cig;
};
#define CNS_WARPS 8
#ifdef HB_KERNELS_ECP
// One WARP per read (hb_cns_read_w): the 512-column pile-up as range updates on a difference array in the warp's shared memory, the majority tests one
// column per lane, lane 0 for the sequential tail (stretch votes between anchors, script).  Reads are handed out dynamically (A.work); the first launch
// (GRAPH = false) is compiled without the graph consensus and queues the reads that need it for the second, whose warps own a graph arena each.
template <bool GRAPH> __global__ void __launch_bounds__(CNS_WARPS * 32) k_ec_cns_w(CnsArgs A)
{
	extern __shared__ uint64_t cns_smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; uint64_t *S = cns_smem + (size_t)warp * HB_CNS_SMEM_WORDS;
	const uint64_t wid = (uint64_t)blockIdx.x * CNS_WARPS + warp;
	RsScratch W; { int32_t *b = A.rs + wid * HB_CNS_RS_WORDS; W.bb = b; W.be = b + 256; W.st = (RsFrame *)(b + 512); }
	const uint64_t n_units = GRAPH ? A.n_queue : A.nR;
	CnsG G; memset(&G, 0, sizeof(G));
	if (GRAPH) {
		G.nd = A.g_nd + wid * A.g_nodes; G.ncap = A.g_nodes; G.arc = A.g_arc + wid * A.g_arcs; G.arc_cap = A.g_arcs; G.q = A.g_q + wid * A.g_nodes; G.q_cap = A.g_nodes;
		G.b32 = A.g_b32 + wid * A.g_arcs; G.b32_cap = A.g_arcs; G.nseq = A.g_ns + wid * A.g_nseq; G.nseq_np = A.g_np + wid * A.g_nseq; G.nseq_cap = A.g_nseq;
		G.ez.path = A.g_path + wid * A.g_pcap; G.ez.pcap = A.g_pcap; G.ez.vec = A.g_vec + wid * 32; G.ez.vstride = 2; G.ez.warp = 0; G.ez.cig = A.g_cig + wid * A.g_ccap; G.ez.ccap = (int32_t)A.g_ccap;
	}
	for (;;) {
		uint32_t u = 0; if (lane == 0) u = atomicAdd(A.work, 1u);
		u = __shfl_sync(0xffffffffu, u, 0);
		if (u >= n_units) break;
		const uint64_t r = GRAPH ? A.queue[u] : u;
		const uint64_t o0 = A.o_off[r], rid = A.r0 + r, e0 = A.ent_off[r]; const uint32_t n = (uint32_t)(A.o_off[r + 1] - o0);
		PhPair *ord = (PhPair *)(A.ord + o0); CnsOv *ov = A.cov + o0; uint32_t n_ov = 0, hdr = 0; // hdr: bit 0-3 status bits, bit 8 = dedup overflow
		if (lane == 0) {
			int ovf = 0; uint8_t st = 0;
			for (uint32_t j = 0; j < n; j++) if (A.ph[o0 + j].st == 2 && A.ph[o0 + j].need_rechain) st |= 4; // the re-seeding rescue of an overlap of this read ran out of scratch: its lists are not final
			const uint32_t keep = hb_ec_dedup(A.ph + o0, n, ord, W, &ovf);
			if (!ovf) for (uint32_t k = 0; k < keep; k++) {
				const hb_phase_t &z = A.ph[o0 + ord[k].idx]; const hb_alnb_t &b = A.alnb[o0 + ord[k].idx];
				if (z.is_match != 1 || !b.w_n) continue;
				CnsOv o; o.w = A.wl + b.w_off; o.wn = b.w_n; o.y_id = z.y_id; o.rev = z.rev; ov[n_ov++] = o;
			}
			hdr = st | (ovf ? 256u : 0u);
		}
		n_ov = __shfl_sync(0xffffffffu, n_ov, 0); hdr = __shfl_sync(0xffffffffu, hdr, 0); __syncwarp();
		const uint8_t st = (uint8_t)(hdr & 0xff);
		if (hdr & 256) { if (lane == 0) { atomicOr(A.err, 128); A.out_n[r] = 0; A.status[r] = st | 2; } continue; }
		CnsCtx C; C.R = A.R; C.q = hb_rd_view(A.R, rid, 0); C.ql = A.R.len[rid]; C.ov = ov; C.pool = A.pool; C.ent = A.ent + e0; C.ct = 0; C.b32 = A.b32 + e0;
		C.out = A.out + A.out_off[r]; C.out_cap = (uint32_t)(A.out_off[r + 1] - A.out_off[r]); C.g = GRAPH ? &G : (CnsG *)0;
		const uint64_t nec = hb_cns_read_w<GRAPH>(C, S, n_ov, A.srt + e0, A.act_a + e0, A.act_b + e0, A.key + e0);
		if (lane == 0) {
			if (C.need_full) { A.out_n[r] = 0; A.status[r] = st | 1 | (C.need_full == 2 ? 8 : 0); } // first launch: queued for the second; second launch: the arena was too small (bit 3)
			else if (C.ovf) { A.out_n[r] = 0; A.status[r] = st | 2; atomicOr(A.err, 256); }
			else { A.out_n[r] = C.out_n; A.status[r] = st; atomicAdd(A.nec, (unsigned long long)nec); }
		}
		__syncwarp();
	}
}
#endif // HB_KERNELS_ECP
#ifdef HB_KERNELS_MAIN
__global__ void k_sc_compact(uint64_t nR, const uint64_t *__restrict__ slot_off, const uint64_t *__restrict__ dense_off, const uint16_t *__restrict__ in, uint16_t *__restrict__ out)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	const uint64_t a = slot_off[r], b = dense_off[r], c = dense_off[r + 1] - b;
	for (uint64_t k = 0; k < c; k++) out[b + k] = in[a + k];
}

// ----------------------------------------------------------------------------
// window alignment: ed_band_cal_semi_64_w_absent_diag
// (Levenshtein_distance.h:3727-3776, ed_core_64 3116-3125), one thread per
// window, the whole band (<= 63 bits) in one 64-bit register pair.
// ----------------------------------------------------------------------------
static __device__ __forceinline__ int ed_nt4(char c)
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': case 'U': case 'u': return 3; default: return 4; }
}
__global__ void k_ed_semi64(uint64_t n_cases, const char *__restrict__ pat, const uint64_t *__restrict__ pat_off, const char *__restrict__ txt, const uint64_t *__restrict__ txt_off,
                            const int32_t *__restrict__ thre_a, const int32_t *__restrict__ abs_a, int32_t *__restrict__ err_o, int32_t *__restrict__ pe_o)
{
	uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_cases) return;
	const char *pstr = pat + pat_off[c], *tstr = txt + txt_off[c];
	int32_t pn = (int32_t)(pat_off[c + 1] - pat_off[c]), tn = (int32_t)(txt_off[c + 1] - txt_off[c]), thre = thre_a[c], abs_diag = abs_a[c];
	uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN, X, D0, HN, HP, mm;
	int32_t bd, i, err = abs_diag, i_bd, tn0 = tn - 1, cut = thre + (thre << 1), best = INT32_MAX, pe = -1, site, ai, uge = INT32_MAX, ch;
	bool dead = pn > tn + cut || tn > pn + cut;
	if (!dead) {
		bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn;
		for (i = 0, mm = 1ULL << abs_diag; i < bd; i++) { Peq[ed_nt4(pstr[i])] |= mm; mm <<= 1; }
		i_bd = (thre << 1) - abs_diag; VN = (1ULL << abs_diag) - 1;
		Peq[4] = 0; mm = 1ULL << (thre << 1);
		for (i = 0; i <= tn0; i++) {
			X = Peq[ed_nt4(tstr[i])] | VN;
			D0 = ((VP + (X & VP)) ^ VP) | X;
			HN = VP & D0; HP = VN | ~(VP | D0);
			X = D0 >> 1;
			VN = X & HP; VP = HN | ~(X | HP);
			if (!(D0 & 1ULL)) { ++err; if (err > cut) { dead = true; break; } }
			if (i == tn0) break;
			Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
			++i_bd; ch = 4;
			if (i_bd < pn) ch = ed_nt4(pstr[i_bd]);
			if (ch < 4) Peq[ch] |= mm;
		}
	}
	if (!dead) {
		site = tn - 1 - abs_diag; ai = pn - tn + abs_diag;
		for (i = 0; site < 0 && i < ai; i++, site++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
		if (err <= thre && err <= best) { best = err; pe = site; }
		site -= i;
		while (i < ai) {
			err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); ++i;
			if (err <= thre && err <= best) { best = err; pe = site + i; }
			if (i == thre) uge = err;
		}
		if (uge <= thre && uge == best) pe = site + thre;
	}
	err_o[c] = best; pe_o[c] = pe;
}
#endif // HB_KERNELS_MAIN

// ---- launchers of the kernels compiled in kern_ecb.cu / kern_ecp.cu ----
enum { HB_K_ECB_PREP, HB_K_ECB_SEG_FAST, HB_K_ECB_SEG, HB_K_ECB_SEG_G, HB_K_ECB_SEG_W, HB_K_ECB_MERGE };
void hb_k_ecaln(int slow, unsigned grid, cudaStream_t st, const EcAlnArgs &A);                 // k_ec_overlap_fast (128 threads) / k_ec_overlap (64)
void hb_k_ecb(int which, unsigned grid, unsigned block, cudaStream_t st, const EcCigArgs &A);
void hb_k_ph(int decide, unsigned grid, cudaStream_t st, const PhArgs &A);                     // PH_WARPS * 32 threads
int hb_k_cns(int graph, unsigned grid, cudaStream_t st, const CnsArgs &A);                     // CNS_WARPS * 32 threads, CNS_WARPS * HB_CNS_SMEM_WORDS * 8 bytes of shared memory
