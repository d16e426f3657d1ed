// hb_sketch.cuh — minimizer sketch of one read.
//
// Behaviour of mz1_ha_sketch (sketch.cpp:454-579) + mz1_select_mz_h
// (sketch.cpp:247-330) + mz1_hf_select (194-216), restructured for the GPU:
// the read is streamed from its 2-bit packed form 32 bases per 64-bit load, the
// (w)-slot candidate ring lives in shared memory (one column per thread), and
// minimizers go straight to the read's slice of the batch output.
#pragma once
#include "hb_common.cuh"

struct SketchPar { int32_t w, k, is_hpc, sample_dist, rewin; };

// candidate = {x, meta}; meta = cnt:28 | pos:27 | rev:1 | span:8 (ha_mz1_t with
// the filter-table count in .rid, sketch.cpp:516)
#define SK_CNT(m) ((uint32_t)((m) & 0xfffffffULL))
#define SK_DUMMY_META 0xfffffffULL
HB_HD int sk_cmp(uint64_t ax, uint64_t am, uint64_t bx, uint64_t bm)
{ // mz1_mzcmp, sketch.cpp:184
	uint32_t ca = SK_CNT(am), cb = SK_CNT(bm);
	if (ca != cb) return ca < cb ? -1 : 1;
	return (ax > bx) - (ax < bx);
}

// Ring accessors: S = stride between consecutive slots of one thread's ring
template <typename T> struct RingRef {
	T *base; int stride;
	HB_HD T &operator[](int j) const { return base[(size_t)j * stride]; }
};

struct SketchOut {
	hb_mz_t *mz; uint32_t *l; uint32_t cap; uint32_t n; int ovf;
	HB_HD void push(uint64_t x, uint64_t meta, uint32_t lv)
	{
		if (n < cap) { mz[n].x = x; mz[n].info = meta; l[n] = lv; }
		else ovf = 1;
		n++;
	}
};

#define SK_MARK 0x80000000u
#define SK_L(o, i) ((int64_t)((o).l[i] & 0x7fffffffu))
#define SK_ACT(o, i) ((i) >= 0 && SK_CNT((o).mz[i].info) > 0)

HB_HD int sk_cmp_l(const SketchOut &o, int32_t ai, int32_t bi)
{ // mz1_mzcmp_l, sketch.cpp:217-225
	if (ai >= 0 && bi >= 0) {
		uint32_t ca = SK_CNT(o.mz[ai].info), cb = SK_CNT(o.mz[bi].info);
		if (ca > 0 && cb > 0) return sk_cmp(o.mz[ai].x, o.mz[ai].info, o.mz[bi].x, o.mz[bi].info);
		return (ca == 0) - (cb == 0);
	}
	return (ai < 0) - (bi < 0);
}

HB_HD void sk_rescan(SketchOut &o, int32_t si, int32_t i, int32_t *mi, int skip_inactive)
{ // "new minimum of [si,i], then mark every equal" (sketch.cpp:232-241, 282-290, 297-305)
	int32_t m;
	for (m = si, *mi = -1; m <= i; m++) {
		if (skip_inactive && !SK_ACT(o, m)) continue;
		if (sk_cmp_l(o, *mi, m) >= 0) *mi = m;
	}
	if (SK_ACT(o, *mi))
		for (m = si; m <= i; m++) {
			if (!SK_ACT(o, m)) continue;
			if (sk_cmp_l(o, *mi, m) == 0) o.l[m] |= SK_MARK;
		}
}

HB_HD int sk_h_lt(const hb_mz_t &a, const hb_mz_t &b) { return sk_cmp(a.x, a.info, b.x, b.info) < 0; }
HB_HD void sk_heapdown(int i, int n, hb_mz_t *l, int32_t *li)
{ // ks_heapdown, ksort.h:43-53
	int k = i; hb_mz_t tmp = l[i]; int32_t ti = li[i];
	while ((k = (k << 1) + 1) < n) {
		if (k != n - 1 && sk_h_lt(l[k], l[k + 1])) ++k;
		if (sk_h_lt(l[k], tmp)) break;
		l[i] = l[k]; li[i] = li[k]; i = k;
	}
	l[i] = tmp; li[i] = ti;
}
HB_HD void sk_hf_select(SketchOut &o, int32_t si, int32_t ei, int32_t n, int32_t len, int32_t sample_dist)
{ // mz1_hf_select, sketch.cpp:194-216: keep the max_high_occ smallest of a streak
	hb_mz_t b[16]; int32_t bi[16];
	int32_t ps, pe, j, k, max_occ;
	if (ei - si <= 1) return;
	ps = si < 0 ? 0 : (int32_t)HB_MZ_POS(o.mz[si].info);
	pe = ei == n ? len : (int32_t)HB_MZ_POS(o.mz[ei].info);
	max_occ = (int32_t)((double)(pe - ps) / sample_dist + .499);
	if (max_occ > 16) max_occ = 16;
	for (j = si + 1, k = 0; j < ei && k < max_occ; ++j, ++k) { b[k] = o.mz[j]; bi[k] = j; }
	for (int i = (k >> 1) - 1; i >= 0; --i) sk_heapdown(i, k, b, bi);
	for (; j < ei; ++j)
		if (sk_h_lt(o.mz[j], b[0])) { b[0] = o.mz[j]; bi[0] = j; sk_heapdown(0, k, b, bi); }
	for (j = 0; j < k; ++j)
		if ((int32_t)SK_CNT(b[j].info) < pe - ps) o.mz[bi[j]].info &= ~0xfffffffULL;
}

HB_HD void sk_select_mz_h(SketchOut &o, int32_t len, int32_t sample_dist, int32_t w, int32_t k, int32_t tot_l)
{ // mz1_select_mz_h, sketch.cpp:247-330 (w = mz_rewin)
	int32_t i, mi = -1, si, last0, n = (int32_t)o.n, m, ws = w + k - 1, any = 0;
	if (n == 0) return;
	for (i = 0, last0 = -1; i <= n; ++i) {
		if (i == n || SK_CNT(o.mz[i].info) == 0) {
			if (i - last0 > 1) {
				int32_t ps = last0 < 0 ? 0 : (int32_t)HB_MZ_POS(o.mz[last0].info);
				int32_t pe = i == n ? len : (int32_t)HB_MZ_POS(o.mz[i].info);
				if (((int32_t)((double)(pe - ps) / sample_dist + .499)) > 0) { any = 1; break; }
			}
			last0 = i;
		}
	}
	if (!any) return;
	for (si = 0, i = 0, mi = -1; i < n; i++) {
		if (SK_L(o, i) >= ws || (i + 1 < n && SK_L(o, i) < ws && SK_L(o, i + 1) > ws) || (i + 1 == n && tot_l >= ws && SK_L(o, i) < ws)) {
			sk_rescan(o, si, i, &mi, 1);
			break;
		}
	}
	if (i < n) {
		for (si = 0, i++; i < n; i++) {
			for (; si < i; si++)
				if (SK_L(o, si) + w > SK_L(o, i)) break;
			if (sk_cmp_l(o, i, mi) <= 0) {
				if (SK_ACT(o, mi)) o.l[mi] |= SK_MARK;
				mi = i;
			} else if (si > mi) {
				if (SK_ACT(o, mi)) o.l[mi] |= SK_MARK;
				sk_rescan(o, si, i, &mi, 0);
			}
		}
		if (SK_ACT(o, mi)) o.l[mi] |= SK_MARK;
		for (i = n - 1; si < n && SK_L(o, si) + w <= tot_l + 1; si++)
			if (si > mi) {
				if (SK_ACT(o, mi)) o.l[mi] |= SK_MARK;
				sk_rescan(o, si, i, &mi, 0);
			}
		for (i = 0, last0 = -1; i <= n; ++i) {
			if (i == n || SK_CNT(o.mz[i].info) == 0) {
				if (i - last0 > 1) {
					int32_t ps = last0 < 0 ? 0 : (int32_t)HB_MZ_POS(o.mz[last0].info);
					int32_t pe = i == n ? len : (int32_t)HB_MZ_POS(o.mz[i].info);
					if (((int32_t)((double)(pe - ps) / sample_dist + .499)) > 0) {
						for (m = last0 + 1, mi = 0; m < i; ++m)
							if (o.l[m] & SK_MARK) { o.mz[m].info &= ~0xfffffffULL; mi++; }
						if (mi == 0) sk_hf_select(o, last0, i, n, len, sample_dist);
					}
				}
				last0 = i;
			}
		}
	}
	for (i = n = 0; i < (int32_t)o.n; ++i)
		if (SK_CNT(o.mz[i].info) == 0) { o.mz[n] = o.mz[i]; n++; }
	o.n = (uint32_t)n;
}


// ===========================================================================
// Two-stage sketch (the production path): the sequential scan of a read only
// produces its stream of ring events; window minima and emissions are then
// decided independently per event.
//
// Why this is exact.  In mz1_ha_sketch (sketch.cpp:521-569) the ring advances
// once per accepted HPC symbol (one whose k-mer is not strand-symmetric) and once
// per N base; `min` is always the right-most minimal entry of the last w ring
// entries (ties go to the newer entry: `>=` at 543 and 555-557; the initial 0xff
// fill compares equal to a dummy).  So with E[t] the t-th ring event and
// M(t) = right-most argmin of E(t-w, t], everything the loop emits at event t is a
// function of M(t-1), M(t), E(t-w..t] and l(t):
//   l == w+k-1 and M(t-1) real      -> entries of E[t-w+1, t-1] equal to M(t-1), other position   (523-534)
//   E[t] <= M(t-1)                  -> M(t-1) if l >= w+k and real                                 (543-547)
//   else if M(t-1) is E[t-w]        -> M(t-1) if l >= w+k-1 and real; then, if M(t) real, the
//                                      entries of E[t-w+1, t] equal to M(t), other position       (548-568)
// and the last M is flushed at the end (571-573).
// ===========================================================================
// one ring event = 16 bytes {x, m}: m = ~0 marks an N base (l restarts), m = SK_DUMMY_META an
// accepted symbol without a candidate k-mer; l is recomputed from the stream (l = events since the last N)
struct SkEv { ulonglong2 *e; };
#define SK_NMARK (~0ULL)

// stage 1: one thread per read; writes the read's ring events (cap = read length + 1).
// The read is streamed 128 bases (one 32-byte sector) per load.  The k-symbol span queue of
// the reference (tiny_queue_t, htab.h:39-57) is replaced by a second cursor into the packed
// read that trails k HPC symbols behind: span = (current run end) - (trailing cursor) + 1.
struct SkChunk { uint64_t w0, w1, w2, w3; int32_t idx; };
HB_HD void sk_chunk_load(SkChunk &c, const uint64_t *seq64, int32_t chunk)
{
#ifdef __CUDA_ARCH__
	const ulonglong2 a = __ldg((const ulonglong2 *)(seq64 + 4 * (size_t)chunk)), b = __ldg((const ulonglong2 *)(seq64 + 4 * (size_t)chunk + 2));
	c.w0 = a.x; c.w1 = a.y; c.w2 = b.x; c.w3 = b.y;
#else
	c.w0 = seq64[4 * (size_t)chunk]; c.w1 = seq64[4 * (size_t)chunk + 1]; c.w2 = seq64[4 * (size_t)chunk + 2]; c.w3 = seq64[4 * (size_t)chunk + 3];
#endif
	c.idx = chunk;
}
HB_HD int sk_chunk_base(SkChunk &c, const uint64_t *seq64, int32_t ii)
{
	if ((ii >> 7) != c.idx) sk_chunk_load(c, seq64, ii >> 7);
	const int q = (ii >> 5) & 3;
	const uint64_t w = q == 0 ? c.w0 : q == 1 ? c.w1 : q == 2 ? c.w2 : c.w3;
	return (int)((w >> ((((ii) & 31) >> 2 << 3) + ((3 - ((ii) & 3)) << 1))) & 3);
}

HB_HD void hb_sketch_events(const DevReads &R, const DevFt &ft, const SketchPar &P, uint64_t rid, SkEv ev, uint32_t *n_ev, uint32_t *tl_out)
{
	const int32_t k = P.k, len = (int32_t)R.len[rid];
	const uint64_t shift1 = k - 1, mask = (1ULL << k) - 1;
	const uint64_t *seq64 = (const uint64_t *)(R.packed + R.off[rid]); // 32-byte aligned
	uint64_t pl0 = 0, pl1 = 0, pl2 = 0, pl3 = 0;
	SkChunk cur, trail; cur.idx = trail.idx = -1; cur.w0 = cur.w1 = cur.w2 = cur.w3 = trail.w0 = trail.w1 = trail.w2 = trail.w3 = 0;
	int32_t i, l = 0, tl = 0, span = 0, qc = 0, tail = 0; uint32_t t = 0;
	uint64_t ni = R.noff[rid], ne = R.noff[rid + 1];
	int32_t next_n = ni < ne ? (int32_t)R.npos[ni] : INT32_MAX;
	for (i = 0; i < len; ++i) {
		int c = sk_chunk_base(cur, seq64, i);
		ulonglong2 e; e.x = ~0ULL; e.y = SK_DUMMY_META;
		if (i == next_n) { c = 4; ++ni; next_n = ni < ne ? (int32_t)R.npos[ni] : INT32_MAX; }
		if (c < 4) {
			int z;
			if (P.is_hpc) { // sketch.cpp:480-492
				const int32_t i0 = i;
				while (i + 1 < len && i + 1 != next_n && sk_chunk_base(cur, seq64, i + 1) == c) ++i;
				if (qc == 0) tail = i0;
				if (++qc > k) { // drop the oldest symbol: step the trailing cursor over its run
					const int c0 = sk_chunk_base(trail, seq64, tail);
					do ++tail; while (sk_chunk_base(trail, seq64, tail) == c0);
					--qc;
				}
				span = i - tail + 1;
			} else span = l + 1 < k ? l + 1 : k;
			pl0 = (pl0 << 1 | (uint64_t)(c & 1)) & mask;
			pl1 = (pl1 << 1 | (uint64_t)(c >> 1)) & mask;
			pl2 = pl2 >> 1 | (uint64_t)(1 - (c & 1)) << shift1;
			pl3 = pl3 >> 1 | (uint64_t)(1 - (c >> 1)) << shift1;
			if (pl1 == pl3) continue;
			z = pl1 < pl3 ? 0 : 1;
			++l; ++tl;
			if (l >= k && span < 256) {
				uint64_t y = z ? hb_hash64(pl2) + hb_hash64(pl3) : hb_hash64(pl0) + hb_hash64(pl1);
				int32_t cnt = hb_ft_lookup(ft, y);
				if (!(cnt >= 1 << 28)) { e.x = y; e.y = (uint64_t)(uint32_t)cnt | (uint64_t)i << 28 | (uint64_t)z << 55 | (uint64_t)span << 56; }
			}
		} else { l = 0; qc = 0; span = 0; e.y = SK_NMARK; }
		ev.e[t] = e; t++;
	}
	*n_ev = t; *tl_out = (uint32_t)tl;
}

#define SK2_POS(m) ((uint32_t)(((m) >> 28) & 0x7ffffffULL))
// right-most argmin of two candidates a (older range) and b (newer range); -1 = empty
#define SK2_PICK(ax, am, ai, bx, bm, bi) (((bi) >= 0 && ((ai) < 0 || sk_cmp((bx), (bm), (ax), (am)) <= 0)) ? 1 : 0)

// stage 2 helpers, written over a window view: X/Mt/L index events by t - base.
// pre[t-base] / suf[t-base]: right-most argmin (absolute t) over [blockstart(t), t] / [t, blockend(t)], blocks of w aligned at 0.
template <typename AX, typename AM>
HB_HD int32_t sk2_window_min(const AX &X, const AM &Mt, const int32_t *pre, const int32_t *suf, int32_t base, int32_t t, int32_t w)
{ // right-most argmin of E(t-w, t]; t < 0 -> -1 (dummy)
	if (t < 0) return -1;
	int32_t lo = t - w + 1;
	if (lo <= 0 || lo % w == 0) return pre[t - base]; // window = one (possibly short) block prefix
	int32_t a = suf[lo - base], b = pre[t - base];
	return sk_cmp(X[b - base], Mt[b - base], X[a - base], Mt[a - base]) <= 0 ? b : a;
}

// emissions of event t; if out != 0 they are written to out[0..), returns their number
template <typename AX, typename AM, typename AL>
// dup_p / dup_c: "some other entry of that window has the same (count, hash) as its minimum"
// (always safe to pass true; false skips the scans for identical minima)
HB_HD uint32_t sk2_emit(const AX &X, const AM &Mt, const AL &L, int32_t base, int32_t t, int32_t mp /*M(t-1)*/, int32_t mc /*M(t)*/, int32_t w, int32_t k,
                        hb_mz_t *out, uint32_t *out_l, bool dup_p = true, bool dup_c = true)
{
	const uint32_t l = L[t - base]; uint32_t n = 0;
	const bool p_real = mp >= 0 && X[mp - base] != ~0ULL;
	uint64_t px = p_real ? X[mp - base] : ~0ULL, pm = p_real ? Mt[mp - base] : SK_DUMMY_META;
	if (mp >= 0 && !p_real) { px = X[mp - base]; pm = Mt[mp - base]; }
	if ((int32_t)l == w + k - 1 && p_real && dup_p) { // sketch.cpp:523-534
		for (int32_t u = t - w + 1 < 0 ? 0 : t - w + 1; u < t; u++)
			if (sk_cmp(px, pm, X[u - base], Mt[u - base]) == 0 && SK2_POS(Mt[u - base]) != SK2_POS(pm)) {
				if (out) { out[n].x = X[u - base]; out[n].info = Mt[u - base]; out_l[n] = L[u - base]; }
				n++;
			}
	}
	if (sk_cmp(px, pm, X[t - base], Mt[t - base]) >= 0) { // sketch.cpp:543-547
		if ((int32_t)l >= w + k && p_real) { if (out) { out[n].x = px; out[n].info = pm; out_l[n] = L[mp - base]; } n++; }
	} else if (mp == t - w) { // sketch.cpp:548-568
		if ((int32_t)l >= w + k - 1 && p_real) { if (out) { out[n].x = px; out[n].info = pm; out_l[n] = L[mp - base]; } n++; }
		const uint64_t cx = X[mc - base], cm = Mt[mc - base];
		if ((int32_t)l >= w + k - 1 && cx != ~0ULL && dup_c)
			for (int32_t u = t - w + 1 < 0 ? 0 : t - w + 1; u <= t; u++)
				if (sk_cmp(cx, cm, X[u - base], Mt[u - base]) == 0 && SK2_POS(cm) != SK2_POS(Mt[u - base])) {
					if (out) { out[n].x = X[u - base]; out[n].info = Mt[u - base]; out_l[n] = L[u - base]; }
					n++;
				}
	}
	return n;
}
