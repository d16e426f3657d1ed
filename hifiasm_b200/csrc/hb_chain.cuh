// hb_chain.cuh — anchor chaining of one (query,target) group and the per-read
// chain post-filter.
//
// Behaviour of lchain_qdp_mcopy_fast / quick_ck_lchain / comput_sc_ch_ec /
// cal_bw / push_ovlp_chain_qgen / gen_fake_cigar (Hash_Table.cpp:2097-2284,
// 2007-2094, 1515-1541, 1475-1488, 1752-1780, 88-109) and of
// lchain_qgen_mcopy_fast's capping / sorting / shadow filter (anchor.cpp:
// 1920-2100), re-laid-out for flat HBM buffers: a group's anchors arrive
// grouped by the probe kernel, are ordered here, chained with int32 DP state in
// a scratch slice, and the chains land in pre-assigned slots of the read.
#pragma once
#include "hb_common.cuh"

struct ChainPar {
	double pen_gap, pen_skip, bw_rate; // chn_pen_gap/skip (host expf, anchor.cpp:2280-2284), bw_thres
	int32_t max_skip, max_iter, max_dis; // 25, 5000, 5000 (anchor.cpp:2276)
	int32_t mcopy_num, mcopy_khit_cutoff; double mcopy_rate; // 3, 32, 0.7 (ecovlp.cpp:3957)
	int32_t max_n_chain; uint32_t chain_cutoff, ocv_w; // asm_opt.max_n_chain, 2, 3072
};

// group directory entry written by the probe/group kernel
#define HB_CHAIN_INPLACE 0x80000000u // hb_chain_t.pad bit: the chain's anchors are a slice of the grouped anchor buffer, not of the chain-anchor buffer
struct GroupDir { uint32_t read, start, count, slot; }; // start: rel. to the read's anchor base; slot: rel. to the read's chain-slot base

#define HB_LINK_FAIL INT32_MIN

HB_HD int32_t hb_link_bw(const hb_hit_t &ai, const hb_hit_t &aj, double bw_rate, int64_t sf_l, int64_t ot_l)
{ // cal_bw, Hash_Table.cpp:1475-1488
	int64_t sf_s = aj.self_offset, sf_e = (int64_t)ai.self_offset + 1;
	int64_t ot_s = aj.offset, ot_e = (int64_t)ai.offset + 1;
	int64_t sf_r = sf_l - sf_e, ot_r = ot_l - ot_e;
	if (sf_s <= ot_s) sf_s = 0; else sf_s -= ot_s;
	if (sf_r <= ot_r) sf_e = sf_l; else sf_e += ot_r;
	return (int32_t)((double)(sf_e - sf_s) * bw_rate);
}

HB_HD int32_t hb_link_sc(const hb_hit_t &ai, const hb_hit_t &aj, const ChainPar &P, int64_t sl, int64_t ol, int32_t *pdd)
{ // comput_sc_ch_ec, Hash_Table.cpp:1515-1541 — IEEE double, no FMA contraction
	int32_t dq, dr, dd, dg, q_span, sc, w;
	dq = (int32_t)((int64_t)ai.self_offset - (int64_t)aj.self_offset);
	if (dq <= 0) return HB_LINK_FAIL;
	dr = (int32_t)((int64_t)ai.offset - (int64_t)aj.offset);
	if (dr <= 0) return HB_LINK_FAIL;
	dd = dr > dq ? dr - dq : dq - dr;
	if (dd > 16 && dd > hb_link_bw(ai, aj, P.bw_rate, sl, ol)) return HB_LINK_FAIL;
	dg = dr < dq ? dr : dq;
	q_span = ai.cnt & 0xffu;
	sc = q_span < dg ? q_span : dg;
	w = (int32_t)(ai.cnt >> 8);
	sc = sc >= w ? sc / w : 1; // normal_w, Hash_Table.cpp:20
	if (dd || (dg > q_span && dg > 0)) {
#ifdef __CUDA_ARCH__
		double lin_pen = __dmul_rn(P.pen_gap, (double)dd);
		double a_pen = __dmul_rn((double)sc, __ddiv_rn(__ddiv_rn((double)dd, (double)dg), P.bw_rate));
		if (dd < 4) lin_pen = lin_pen > a_pen ? a_pen : lin_pen;
		else lin_pen = lin_pen < a_pen ? a_pen : lin_pen;
		lin_pen = __dadd_rn(lin_pen, __dmul_rn(P.pen_skip, (double)dg));
#else
		double lin_pen = P.pen_gap * (double)dd, a_pen = ((double)sc) * ((((double)dd) / ((double)dg)) / P.bw_rate);
		if (dd < 4) lin_pen = lin_pen > a_pen ? a_pen : lin_pen;
		else lin_pen = lin_pen < a_pen ? a_pen : lin_pen;
		lin_pen += P.pen_skip * (double)dg;
#endif
		sc -= (int32_t)lin_pen;
	}
	if (pdd) *pdd = dd;
	return sc;
}

HB_HD void hb_push_chain(hb_chain_t &o, int64_t xl, int64_t yl, int64_t sc, const hb_hit_t &beg, const hb_hit_t &end)
{ // push_ovlp_chain_qgen, Hash_Table.cpp:1752-1780
	int64_t xr, yr;
	o.y_id = HB_HIT_ID(beg); o.y_pos_strand = HB_HIT_ST(beg);
	o.x_pos_s = beg.self_offset; o.y_pos_s = beg.offset;
	o.x_pos_e = end.self_offset; o.y_pos_e = end.offset;
	if (o.x_pos_s <= o.y_pos_s) { o.y_pos_s -= o.x_pos_s; o.x_pos_s = 0; }
	else { o.x_pos_s -= o.y_pos_s; o.y_pos_s = 0; }
	xr = xl - o.x_pos_e - 1; yr = yl - o.y_pos_e - 1;
	if (xr <= yr) { o.x_pos_e = (uint32_t)(xl - 1); o.y_pos_e += (uint32_t)xr; }
	else { o.y_pos_e = (uint32_t)(yl - 1); o.x_pos_e += (uint32_t)yr; }
	o.shared_seed = (int32_t)sc; o.n_hits = 0; o.first_hit = 0; o.fc_off = 0; o.fc_n = 0; o.pad = 0;
}

// Fake_Cigar writer (add_fake_cigar, Hash_Table.cpp:1295-1328)
struct FcOut { uint64_t *buf; uint32_t n, cap; int ovf;
	HB_HD void add(uint32_t site, int32_t shift)
	{
		uint32_t e = shift < 0 ? ((uint32_t)(-shift) << 1 | 1u) : (uint32_t)shift << 1;
		if (n < cap) buf[n] = (uint64_t)site << 32 | e; else ovf = 1;
		n++;
	}
};
HB_HD int hb_fc_shift(uint64_t e) { uint32_t t = (uint32_t)e; int r = (int)(t >> 1); return (t & 1) ? -r : r; }

HB_HD void hb_gen_fcigar(FcOut &fc, hb_chain_t &o, const hb_hit_t *hit, int64_t n_hit)
{ // gen_fake_cigar with apend_be = 1, Hash_Table.cpp:88-109
	int64_t k, dq, dr, dd, pdd = INT32_MAX;
	if (fc.buf == 0) return;
	o.fc_off = fc.n;
	fc.add(o.x_pos_s, 0);
	for (k = 0; k < n_hit; k++) {
		dq = (uint32_t)(hit[k].self_offset - o.x_pos_s);
		dr = (uint32_t)(hit[k].offset - o.y_pos_s);
		dd = dr - dq;
		if (dd != pdd) { pdd = dd; fc.add(hit[k].self_offset, (int32_t)pdd); }
	}
	if (!fc.ovf && (int64_t)(fc.buf[fc.n - 1] >> 32) != (int64_t)o.x_pos_e) fc.add(o.x_pos_e, hb_fc_shift(fc.buf[fc.n - 1]));
	o.fc_n = fc.n - o.fc_off;
}

HB_HD void hb_heapsort_i64(int64_t *a, int64_t n)
{ // any correct ascending sort reproduces radix_sort_hc64i (keys are distinct)
	int64_t i, k, c; int64_t tmp;
	for (i = (n >> 1) - 1; i >= 0; --i) {
		tmp = a[i];
		for (k = i; (c = (k << 1) + 1) < n; k = c) { if (c + 1 < n && a[c] < a[c + 1]) ++c; if (a[c] <= tmp) break; a[k] = a[c]; }
		a[k] = tmp;
	}
	for (i = n - 1; i > 0; --i) {
		tmp = a[i]; a[i] = a[0];
		for (k = 0; (c = (k << 1) + 1) < i; k = c) { if (c + 1 < i && a[c] < a[c + 1]) ++c; if (a[c] <= tmp) break; a[k] = a[c]; }
		a[k] = tmp;
	}
}

// order a (nearly ordered) group: strand, then self_offset, then offset
HB_HD void hb_order_group(hb_hit_t *a, int32_t a_n)
{
	for (int32_t i = 1; i < a_n; i++) {
		hb_hit_t v = a[i]; int32_t j;
		uint64_t kv = (uint64_t)(v.id_strand >> 31) << 63 | (uint64_t)v.self_offset << 32 | v.offset;
		for (j = i; j > 0; --j) {
			const hb_hit_t &u = a[j - 1];
			uint64_t ku = (uint64_t)(u.id_strand >> 31) << 63 | (uint64_t)u.self_offset << 32 | u.offset;
			if (ku <= kv) break;
			a[j] = u;
		}
		if (j != i) a[j] = v;
	}
}

struct ChainState { int64_t plus, msc, msc_i, movl, si, ei; };

// quick_ck_lchain, Hash_Table.cpp:2007-2094: strand blocks whose anchors are
// already co-linear are resolved by a linear scan; [si,ei) is left for the DP
HB_HD void hb_chain_quick(const hb_hit_t *a, int32_t a_n, int32_t *f, int32_t *p, int64_t *t, int32_t *ii, const ChainPar &P, int64_t xl, int64_t yl, ChainState &S)
{
	int64_t k, sc, plus, msc, msc_i, movl, si, ei;
	{
		int64_t l, z, sorted = 1;
		plus = 0; msc = msc_i = INT32_MIN; movl = INT32_MAX; si = 0; ei = a_n;
		for (k = 1, l = 0; k <= a_n; k++) {
			if (k == a_n || HB_HIT_ST(a[k]) != HB_HIT_ST(a[l])) {
				t[k - 1] = 0; ii[k - 1] = 0;
				if (sorted) {
					int64_t plus0 = 0, msc0 = INT32_MIN, msc_i0 = INT32_MIN, movl0, ddt = 0;
					p[l] = -1; f[l] = a[l].cnt & 0xffu;
					if (f[l] >= msc0) { msc0 = f[l]; msc_i0 = l; }
					if (f[l] < plus0) plus0 = f[l];
					for (z = l + 1; z < k; z++) {
						int32_t dd = 0, s = hb_link_sc(a[z], a[z - 1], P, xl, yl, &dd); int64_t csc;
						if (s == HB_LINK_FAIL) break;
						sc = (int64_t)s + f[z - 1]; csc = a[z].cnt & 0xffu;
						if (sc < csc) break;
						p[z] = (int32_t)(z - 1); f[z] = (int32_t)sc; ddt += dd;
						if (f[z] >= msc0) { msc0 = f[z]; msc_i0 = z; }
						if (f[z] < plus0) plus0 = f[z];
					}
					if (z >= k && msc_i0 == k - 1) {
						if (k - l >= 2 && ddt > 16 && ddt > hb_link_bw(a[k - 1], a[l], P.bw_rate, xl, yl)) msc_i0 = INT32_MIN;
						if (msc_i0 == k - 1) {
							if (msc0 >= msc) {
								movl0 = hb_chain_len(a[msc_i0].self_offset, a[msc_i0].self_offset, xl, a[msc_i0].offset, a[msc_i0].offset, yl);
								if (msc0 > msc || movl0 < movl) { msc = msc0; msc_i = msc_i0; movl = movl0; }
							}
							if (plus0 < plus) plus = plus0;
							if (ei > k) si = k; else ei = l;
						}
					}
				}
				l = k; sorted = 1;
			} else {
				if (a[k].self_offset <= a[k - 1].self_offset || a[k].offset <= a[k - 1].offset) sorted = 0;
				t[k - 1] = 0; ii[k - 1] = 0;
			}
		}
	}
	S.plus = plus; S.msc = msc; S.msc_i = msc_i; S.movl = movl; S.si = si; S.ei = ei;
}


// backtrack, secondary chains (mcopy), emission: Hash_Table.cpp:2178-2283
HB_HD int32_t hb_chain_finish(const hb_hit_t *a, int32_t a_n, hb_hit_t *des, uint64_t des_abs, int32_t *f, int32_t *p, int64_t *t, int32_t *ii,
                              const ChainPar &P, int64_t xl, int64_t yl, hb_chain_t *out, int32_t n_slots, FcOut &fc, const ChainState &S)
{
	int64_t sc, msc = S.msc, msc_i = S.msc_i, plus = S.plus, min_sc, ch_n, i, k, j, cL = 0;
	int32_t n_out = 0;
	(void)n_slots;
	for (i = msc_i, cL = 0; i >= 0; i = p[i]) { ii[i] = 1; t[cL++] = i; } // Hash_Table.cpp:2178

	if (P.mcopy_num > 1 && cL >= P.mcopy_khit_cutoff) { // Hash_Table.cpp:2180-2270
		msc -= plus; min_sc = (int64_t)((double)msc * P.mcopy_rate); ii[msc_i] = 0;
		for (i = ch_n = 0; i < a_n; ++i) {
			f[i] -= (int32_t)plus; t[i] = 0;
			if (!ii[i] && f[i] >= min_sc) { t[ch_n] = (int64_t)(((uint64_t)(uint32_t)f[i]) << 32) + (i << 1); ch_n++; }
		}
		if (ch_n > 1) {
			int64_t n_v, n_v0, ni, n_u, done = 0;
			hb_heapsort_i64(t, ch_n);
			for (k = ch_n - 1, n_v = n_u = 0; k >= 0 && n_u < P.mcopy_num; --k) {
				n_v0 = n_v;
				for (i = ((uint32_t)t[k]) >> 1; i >= 0 && (t[i] & 1) == 0;) { ii[n_v++] = (int32_t)i; t[i] |= 1; i = p[i]; }
				if (n_v0 == n_v) continue;
				sc = i < 0 ? (t[k] >> 32) : ((t[k] >> 32) - f[i]);
				if (sc >= min_sc) {
					hb_chain_t &z = out[n_out];
					hb_push_chain(z, xl, yl, sc + plus, a[ii[n_v - 1]], a[ii[n_v0]]);
					if (!n_u || n_v - n_v0 > 1) { z.n_hits = (uint32_t)(n_v - n_v0); z.first_hit = (uint32_t)n_v0; n_u++; n_out++; }
					else n_v = n_v0;
				} else n_v = n_v0;
			}
			for (k = 0; k < n_out; k++) { // Hash_Table.cpp:2230-2263
				hb_chain_t &z = out[k];
				n_v0 = z.first_hit; ni = z.n_hits;
				z.first_hit = (uint32_t)(des_abs + done);
				for (j = 0; j < ni; j++) des[done + j] = a[ii[n_v0 + (ni - j - 1)]];
				hb_gen_fcigar(fc, z, des + done, ni);
				done += ni;
			}
			return (int32_t)done;
		}
		msc += plus; i = msc_i; cL = 0;
		while (i >= 0) { t[cL++] = i; i = p[i]; }
	}
	{ // Hash_Table.cpp:2277-2283
		hb_chain_t &z = out[0];
		hb_push_chain(z, xl, yl, msc, a[t[cL - 1]], a[t[0]]);
		for (i = 0; i < cL; i++) des[i] = a[t[cL - i - 1]];
		z.first_hit = (uint32_t)des_abs; z.n_hits = (uint32_t)cL;
		hb_gen_fcigar(fc, z, des, cL);
	}
	return (int32_t)cL;
}


// ---------------------------------------------------------------------------
// klib sorts restated on index arrays: the comparisons read keys through the
// index, the swaps move indices — the same permutation as sorting the structs
// (ksort.h:80-160 introsort/combsort/insertion sort; 172-221 MSD radix).
// ---------------------------------------------------------------------------
struct LtSS { const hb_chain_t *c; HB_HD bool operator()(uint32_t a, uint32_t b) const { return c[a].shared_seed > c[b].shared_seed; } }; // oreg_ss_lt, anchor.cpp:35
struct LtXS { const hb_chain_t *c; HB_HD bool operator()(uint32_t a, uint32_t b) const
	{ return ((uint64_t)c[a].x_pos_s << 32 | c[a].x_pos_e) < ((uint64_t)c[b].x_pos_s << 32 | c[b].x_pos_e); } }; // oreg_xs_lt, anchor.cpp:32

template <typename T, typename LT> HB_HD void hb_insertsort(T *s, T *t, LT lt)
{ T *i, *j, sw; for (i = s + 1; i < t; ++i) for (j = i; j > s && lt(*j, *(j - 1)); --j) { sw = *j; *j = *(j - 1); *(j - 1) = sw; } }
template <typename T, typename LT> HB_HD void hb_combsort(int64_t n, T *a, LT lt)
{ // ks_combsort, ksort.h:88-109
	const double shrink = 1.2473309501039786540366528676643;
	int do_swap; int64_t gap = n; T tmp, *i, *j;
	do {
		if (gap > 2) { gap = (int64_t)((double)gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
		do_swap = 0;
		for (i = a; i < a + n - gap; ++i) { j = i + gap; if (lt(*j, *i)) { tmp = *i; *i = *j; *j = tmp; do_swap = 1; } }
	} while (do_swap || gap > 2);
	if (gap != 1) hb_insertsort(a, a + n, lt);
}
template <typename T, typename LT> HB_HD void hb_introsort(int64_t n, T *a, LT lt)
{ // ks_introsort, ksort.h:110-160
	int d; struct { T *l, *r; int d; } stack[130], *top = stack;
	T rp, sw, *s, *t, *i, *j, *k;
	if (n < 1) return;
	if (n == 2) { if (lt(a[1], a[0])) { sw = a[0]; a[0] = a[1]; a[1] = sw; } return; }
	for (d = 2; (1ll << d) < n; ++d) {}
	s = a; t = a + (n - 1); d <<= 1;
	while (1) {
		if (s < t) {
			if (--d == 0) { hb_combsort(t - s + 1, s, lt); t = s; continue; }
			i = s; j = t; k = i + ((j - i) >> 1) + 1;
			if (lt(*k, *i)) { if (lt(*k, *j)) k = j; }
			else k = lt(*j, *i) ? i : j;
			rp = *k;
			if (k != t) { sw = *k; *k = *t; *t = sw; }
			for (;;) {
				do ++i; while (lt(*i, rp));
				do --j; while (i <= j && lt(rp, *j));
				if (j <= i) break;
				sw = *i; *i = *j; *j = sw;
			}
			sw = *i; *i = *t; *t = sw;
			if (i - s > t - i) {
				if (i - s > 16) { top->l = s; top->r = i - 1; top->d = d; ++top; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { top->l = i + 1; top->r = t; top->d = d; ++top; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == stack) { hb_insertsort(a, a + n, lt); return; }
			--top; s = top->l; t = top->r; d = top->d;
		}
	}
}

// radix_sort_overlap_region_sort on y_id (Hash_Table.cpp:12; ksort.h:172-221),
// in-place MSD radix with 8-bit digits, insertion sort below 65 elements —
// unstable, so it is restated move for move.  KEY(x) reads the 32-bit key.
template <typename T, typename KEY> HB_HD void hb_rs_insertsort(T *beg, T *end, KEY key)
{
	for (T *i = beg + 1; i < end; ++i)
		if (key(*i) < key(*(i - 1))) {
			T *j, tmp = *i;
			for (j = i; j > beg && key(tmp) < key(*(j - 1)); --j) *j = *(j - 1);
			*j = tmp;
		}
}
#define HB_RS_STACK 320
struct RsFrame { int32_t b, e, s; };
struct RsScratch { int32_t *bb, *be; RsFrame *st; }; // 256 + 256 ints and HB_RS_STACK frames
// top_shift: the highest digit level at which the keys can differ (24 = all of them).  A level whose digit is the same for every key leaves one bucket,
// moves nothing and recurses into the whole range one level down, so starting below it is the same sequence of moves.
template <typename T, typename KEY> HB_HD int hb_rs_sort32(T *beg0, T *end0, KEY key, const RsScratch &W, int top_shift = 24)
{ // returns 1 if the explicit stack overflowed (more than 65*HB_RS_STACK elements)
	if (end0 - beg0 <= 64) { hb_rs_insertsort(beg0, end0, key); return 0; }
	// explicit recursion stack of (range, shift): pending ranges are disjoint and
	// each holds > 64 elements, so n <= 65*HB_RS_STACK never overflows it
	RsFrame *st = W.st; int32_t *bb = W.bb, *be = W.be; int sp = 0;
	st[sp].b = 0; st[sp].e = (int32_t)(end0 - beg0); st[sp].s = top_shift; sp++;
	while (sp) {
		RsFrame fr = st[--sp]; T *beg = beg0 + fr.b, *end = beg0 + fr.e; int s = fr.s;
		int k;
		for (k = 0; k < 256; ++k) bb[k] = be[k] = 0;
		for (T *i = beg; i != end; ++i) ++be[key(*i) >> s & 255];
		for (k = 1; k < 256; ++k) { be[k] += be[k - 1]; bb[k] = be[k - 1]; }
		for (k = 0; k < 256;) {
			if (bb[k] != be[k]) {
				int l = key(beg[bb[k]]) >> s & 255;
				if (l != k) {
					T tmp = beg[bb[k]], sw;
					do { sw = tmp; tmp = beg[bb[l]]; beg[bb[l]++] = sw; l = key(tmp) >> s & 255; } while (l != k);
					beg[bb[k]++] = tmp;
				} else ++bb[k];
			} else ++k;
		}
		if (s) {
			int ns = s > 8 ? s - 8 : 0;
			// the reference recurses bucket 0..255 in order; each call is
			// independent of the others, so the order of processing is free
			for (k = 255; k >= 0; --k) {
				int32_t lo = k ? be[k - 1] : 0, hi = be[k];
				if (hi - lo > 64) { if (sp >= HB_RS_STACK) return 1; st[sp].b = fr.b + lo; st[sp].e = fr.b + hi; st[sp].s = ns; sp++; }
				else if (hi - lo > 1) hb_rs_insertsort(beg + lo, beg + hi, key);
			}
		}
	}
	return 0;
}

HB_HD int hb_ov_type(const hb_chain_t &r, uint64_t len)
{ // ha_ov_type, anchor.cpp:86-91
	if (r.x_pos_s == 0 && r.x_pos_e == len - 1) return 2;
	if (r.x_pos_s > 0 && r.x_pos_e < len - 1) return 3;
	return r.x_pos_s == 0 ? 0 : 1;
}

HB_HD void hb_cov_add(uint64_t *cc, uint64_t cwn, uint64_t ocv_w, uint64_t rl, const hb_chain_t &o)
{ // anchor.cpp:1986-1999 / 2028-2041
	uint64_t m = o.x_pos_s / ocv_w, rs = o.x_pos_s, re = (uint64_t)o.x_pos_e + 1, cws, cwe, os, oe;
	for (cws = m * ocv_w; m < cwn; m++) {
		cwe = cws + ocv_w; if (cwe > rl) cwe = rl;
		os = rs >= cws ? rs : cws; oe = re <= cwe ? re : cwe;
		if (oe <= os) break;
		if (((uint32_t)cc[m]) + (oe - os) < UINT32_MAX) cc[m] += oe - os;
		else { cc[m] >>= 32; cc[m] <<= 32; cc[m] |= UINT32_MAX; }
		cws += ocv_w;
	}
}

// Per-read chain post-filter: lchain_qgen_mcopy_fast after the per-target loop
// (anchor.cpp:1944-2100).  ch[0..n_slots): the read's chain slots in target
// order (n_hits == 0 = empty slot); chits: chain-anchor buffer (absolute
// first_hit); idx: n_slots scratch; cc: coverage-window scratch (rl/ocv_w+1).
// Leaves the kept chains' slot numbers in idx[0..n), returns n; writes the
// compacted anchor index (cl->list position, Hash_Table.cpp:2280) to .pad.
HB_HD uint32_t hb_chain_post(hb_chain_t *ch, uint32_t n_slots, const hb_hit_t *chits_, const hb_hit_t *ghits, uint32_t *idx, uint64_t *cc, uint64_t rl, const ChainPar &P)
{
	uint64_t i, k, l, n = 0, lch = 0, cwn = 0, m = 0, max_n_chain = (uint64_t)P.max_n_chain; uint32_t t;
	for (i = 0; i < n_slots; i++)
		if (ch[i].n_hits) { idx[n++] = (uint32_t)i; ch[i].pad = (ch[i].pad & HB_CHAIN_INPLACE) | (uint32_t)m; m += ch[i].n_hits; if (ch[i].n_hits < P.chain_cutoff) lch = 1; }
	if (P.chain_cutoff < 2) lch = 0;
	k = n;
	if (n > max_n_chain) { // anchor.cpp:1954-2058
		int32_t w, nn[4] = { 0, 0, 0, 0 }, s[4] = { 0, 0, 0, 0 };
		LtSS lt = { ch };
		hb_introsort((int64_t)n, idx, lt);
		for (i = 0; i < n; ++i) {
			w = hb_ov_type(ch[idx[i]], rl); ++nn[w];
			if ((uint64_t)nn[w] == max_n_chain) s[w] = ch[idx[i]].shared_seed;
		}
		if (s[0] > 0 || s[1] > 0 || s[2] > 0 || s[3] > 0) {
			if ((uint64_t)nn[3] >= max_n_chain && rl >= P.ocv_w) {
				uint64_t cws, cwe;
				cwn = rl / P.ocv_w + (rl % P.ocv_w ? 1 : 0);
				for (i = cws = 0; i < cwn; i++) {
					cwe = cws + P.ocv_w; if (cwe > rl) cwe = rl;
					cc[i] = (cwe - cws) * (max_n_chain >> 1);
					if (cc[i] > UINT32_MAX) cc[i] = UINT32_MAX;
					cc[i] <<= 32;
					cws += P.ocv_w;
				}
			}
			for (i = 0, k = 0, lch = 0; i < n; ++i) {
				const hb_chain_t &o = ch[idx[i]];
				w = hb_ov_type(o, rl);
				if (o.shared_seed >= s[w]) {
					if (cwn) hb_cov_add(cc, cwn, P.ocv_w, rl, o);
					if (k != i) { t = idx[k]; idx[k] = idx[i]; idx[i] = t; }
					if (ch[idx[k]].n_hits < P.chain_cutoff) lch = 1;
					++k;
				} else if (w == 3 && cwn > 0) {
					uint64_t mm = o.x_pos_s / P.ocv_w, cw0 = 0, cw1 = 0, rs = o.x_pos_s, re = (uint64_t)o.x_pos_e + 1, cws, cwe, os, oe;
					for (cws = mm * P.ocv_w; mm < cwn; mm++) {
						cwe = cws + P.ocv_w; if (cwe > rl) cwe = rl;
						os = rs >= cws ? rs : cws; oe = re <= cwe ? re : cwe;
						if (oe <= os) break;
						if ((oe - os) + ((uint64_t)((uint32_t)cc[mm])) >= (cc[mm] >> 32)) cw1 += oe - os; else cw0 += oe - os;
						cws += P.ocv_w;
					}
					if ((double)cw0 >= ((double)(cw0 + cw1) * 0.7)) {
						hb_cov_add(cc, cwn, P.ocv_w, rl, o);
						if (k != i) { t = idx[k]; idx[k] = idx[i]; idx[i] = t; }
						if (ch[idx[k]].n_hits < P.chain_cutoff) lch = 1;
						++k;
					}
				}
			}
			n = k;
		}
	}
	{ LtXS lt = { ch }; hb_introsort((int64_t)n, idx, lt); } // anchor.cpp:2060
	if (lch) { // anchor.cpp:2061-2096
		uint64_t zs, ze, rs, re, ob, os, oe, ocn, kn, ms, me, mm, hh; int64_t osc;
		for (i = l = 0; i < n; ++i) {
			const hb_chain_t &zi = ch[idx[i]];
			if (zi.n_hits < P.chain_cutoff) {
				zs = zi.x_pos_s; ze = (uint64_t)zi.x_pos_e + 1;
				ob = (uint64_t)((double)(ze - zs) * 0.95); if (ob < 16) ob = 16; // OFL
				osc = (int64_t)zi.shared_seed * 16;                                // CH_SC
				ocn = (uint64_t)zi.n_hits << 4;                                    // CH_OCC
				for (k = 0; k < n && ze > ch[idx[k]].x_pos_s; k++) {
					const hb_chain_t &zk = ch[idx[k]];
					if (zk.n_hits < P.chain_cutoff) continue;
					if (zk.n_hits < ocn) continue;
					if (zk.shared_seed < osc) continue;
					rs = zk.x_pos_s; re = (uint64_t)zk.x_pos_e + 1;
					os = rs >= zs ? rs : zs; oe = re <= ze ? re : ze;
					if (oe > os && oe - os >= ob) {
						// count the chain's anchors lying inside [os,oe] (anchor.cpp:2077-2083).  The
						// anchors are ordered by self_offset, so the count is taken over the slice
						// [first anchor ending at or after os, last anchor ending at or before oe]
						// instead of the whole chain: same count, a handful of loads.
						mm = zk.first_hit; kn = 0;
						{
							const hb_hit_t *chits = (zk.pad & HB_CHAIN_INPLACE) ? ghits : chits_;
							uint64_t lo_ = 0, hi_ = zk.n_hits;
							while (lo_ < hi_) { uint64_t mid = (lo_ + hi_) >> 1; if (chits[mm + mid].self_offset < os) lo_ = mid + 1; else hi_ = mid; }
							// ... plus the chain's first anchors when they start at the read's first bases: the reference's `ms = me - span` is unsigned
							// and wraps below zero there (self_offset + 1 == span gives ms = 2^64 - 1), so such an anchor passes `ms >= os` whatever os is
							for (hh = 0; hh < lo_ && kn < ocn; hh++) { me = chits[mm + hh].self_offset; if (me >= 256) break; if (me < (chits[mm + hh].cnt & 0xffu)) kn++; }
							for (hh = lo_; hh < zk.n_hits && kn < ocn; hh++) {
								me = chits[mm + hh].self_offset; if (me > oe) break;
								ms = me - (chits[mm + hh].cnt & 0xffu);
								if (ms >= os) kn++;
							}
						}
						if (kn >= ocn) break;
					}
				}
				if (k < n && ze > ch[idx[k]].x_pos_s) continue;
			}
			if (l != i) { t = idx[l]; idx[l] = idx[i]; idx[i] = t; }
			l++;
		}
		n = l;
	}
	return (uint32_t)n;
}
