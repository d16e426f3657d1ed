// hb_ecrechain.cuh — the re-seeding rescue of step B: rechain_aln_hc (Correct.cpp:17669-17748).
//
// After hc_ovlp_base_direct, an unaligned window of the overlap with >= FORCE_SIN_L (512) bases on BOTH reads is
// re-seeded: the exact runs (>= 10 bases) of step A's window alignments inside it become k-mer hits (gen_win_chain
// 17012-17075: extract_exact_cigar 16805-16845, push_khit 16779-16803), they are chained with the end points of the window
// fixed (lchain_qdp_fix, Hash_Table.cpp:2293-2440; lchain_qcheck 1424-1473; comput_sc_ff 1543-1567), the chain's hits are
// split into their two end points and chained once more (gen_single_khit, Correct.cpp:16865-17008; lchain_simple,
// Hash_Table.cpp:2544-2604), and hc_ovlp_base_direct (17425-17515, pre_mode = the window's mode) aligns the pieces between
// those points: the window shrinks to what really differs.  New windows are appended to the overlap's list (one that
// covers the whole old window replaces it), replaced windows are dropped, and the list is sorted by x_start again
// (gen_hc_fast_cigar0, 17843-17866).
//
// One thread per overlap (a rare path: a structural difference of >= 512 bp between two reads of an accepted overlap).
// The reference mutates step A's windows while it re-seeds (gen_backtrace_adv_exz traces a window that has no cigar yet),
// and later estimates read the mutated windows — so the thread works on a private copy of the overlap's step-A list.
#pragma once
#include "hb_ecaln.cuh"
#include "hb_chain.cuh"

struct EcRc {
	DevReads R;
	hb_wl_t *zw; int32_t zcap;                                   // private copy of step A's window list of the overlap
	const uint16_t *poolA;                                       // step A's cigar pool (windows step A traced itself)
	uint16_t *zc; unsigned long long *zc_used; uint64_t zc_cap;  // cigars of the windows traced here (cidx | 0x80000000)
	uint64_t *path1; uint16_t *cig1;                             // scratch of the one-word traced aligner (5 * w_l words, HB_EC_CIG_TMP runs)
	hb_hit_t *h; int32_t hcap, hn; int64_t *t, *p; int32_t *f;   // the re-seeded hits (cl->list beyond cl->length) + chaining state
	double pen_gap, pen_skip; int32_t h_khit;                    // set_lchain_dp_op(is_accurate = 1, mz_k = E_KHIT), host expf
	RsScratch rs;
	int *err; int ovf;
};
#define HB_RC_PRIV 0x80000000u
#define HB_E_KHIT 31        // E_KHIT, ecovlp.cpp:10

HB_HD void hb_rc_put(EcRc &S, int32_t xs, int32_t ys, uint32_t cnt)
{
	if (S.hn >= S.hcap) { S.ovf = 1; return; }
	hb_hit_t &z = S.h[S.hn++]; z.id_strand = 0; z.self_offset = (uint32_t)xs; z.offset = (uint32_t)ys; z.cnt = cnt;
}
HB_HD void hb_rc_push_khit(EcRc &S, int32_t xs, int32_t ys, uint32_t len, const uint32_t *ic)
{ // push_khit, Correct.cpp:16779-16803: a run longer than 255 becomes several hits at the SAME end point
	uint32_t c = (len >= (uint32_t)S.h_khit) ? 1 : 2; if (ic) c = *ic; c <<= 8;
	if (len > 0) {
		while (len >= 0xffu) { hb_rc_put(S, xs, ys, c + 0xffu); len -= 0xffu; }
		if (len) hb_rc_put(S, xs, ys, c + len);
	} else hb_rc_put(S, xs, ys, c + len);
}
HB_HD uint32_t hb_rc_extract(EcRc &S, const uint16_t *cg, uint32_t cn, int32_t ps, int32_t ts, int32_t pmin, int32_t pmax, int32_t tmin, int32_t tmax, int32_t minl, int64_t min_w_l)
{ // extract_exact_cigar, Correct.cpp:16805-16845 (p = target, t = query)
	uint32_t ci = 0, cl, occ = 0, c; int32_t pi = ps, ti = ts, p[2], t[2], poff = -1, toff = -1, maxl = -1, pos, poe, tos, toe, l;
	while (ci < cn && pi < pmax && ti < tmax) {
		c = cg[ci] >> 14; cl = cg[ci] & 0x3fff;
		for (ci++; ci < cn && (uint32_t)(cg[ci] >> 14) == c; ci++) cl += cg[ci] & 0x3fff;
		if (c == 0) {
			p[0] = pi; p[1] = pi + (int32_t)cl; t[0] = ti; t[1] = ti + (int32_t)cl;
			pos = p[0] > pmin ? p[0] : pmin; poe = p[1] < pmax ? p[1] : pmax;
			tos = t[0] > tmin ? t[0] : tmin; toe = t[1] < tmax ? t[1] : tmax;
			if (poe > pos && toe > tos) {
				l = poe - pos;
				if (l == toe - tos) {
					poe--; toe--;
					if (l > maxl) { poff = poe; toff = toe; maxl = l; }
					if (l >= minl) { hb_rc_push_khit(S, toe, poe, (uint32_t)l, 0); occ++; }
				}
			}
			pi += (int32_t)cl; ti += (int32_t)cl;
		} else if (c == 1) { pi += (int32_t)cl; ti += (int32_t)cl; }
		else if (c == 2) pi += (int32_t)cl;
		else ti += (int32_t)cl;
	}
	if (maxl > 0 && maxl < minl && ts >= tmin && (int64_t)ti >= min_w_l + ts) { const uint32_t w = 3; hb_rc_push_khit(S, toff, poff, (uint32_t)maxl, &w); occ++; }
	return occ;
}

HB_HD int32_t hb_rc_normal_w(int32_t x, int32_t y) { return x >= y ? x / y : 1; } // normal_w, Hash_Table.cpp:20
HB_HD int32_t hb_rc_sc_ff(const hb_hit_t &ai, const hb_hit_t &aj, double bw_rate, double pen_gap, double pen_skip)
{ // comput_sc_ff, Hash_Table.cpp:1543-1567 — IEEE double, no FMA contraction
	int32_t dq, dr, dd, dg, q_span, sc;
	const int64_t dq64 = (int64_t)ai.self_offset - (int64_t)aj.self_offset, dr64 = (int64_t)ai.offset - (int64_t)aj.offset;
	dq = (int32_t)dq64; if (dq < 0) return INT32_MIN;
	dr = (int32_t)dr64; if (dr < 0) return INT32_MIN;
	dd = dr > dq ? dr - dq : dq - dr;
	dg = dr < dq ? dr : dq;
	q_span = (int32_t)(ai.cnt & 0xffu);
	sc = q_span < dg ? q_span : dg;
	sc = hb_rc_normal_w(sc, (int32_t)(ai.cnt >> 8));
	if (dd || (dg > q_span && dg > 0)) {
#ifdef __CUDA_ARCH__
		double lin_pen = __dmul_rn(pen_gap, (double)dd);
		const double a_pen = __dmul_rn((double)sc, __ddiv_rn(__ddiv_rn((double)dd, (double)dg), bw_rate));
		if (lin_pen > a_pen) lin_pen = a_pen;
		lin_pen = __dadd_rn(lin_pen, __dmul_rn(pen_skip, (double)dg));
#else
		double lin_pen = pen_gap * (double)dd; const double a_pen = ((double)sc) * ((((double)dd) / ((double)dg)) / bw_rate);
		if (lin_pen > a_pen) lin_pen = a_pen;
		lin_pen += pen_skip * (double)dg;
#endif
		sc -= (int32_t)lin_pen;
	}
	return sc;
}
HB_HD int32_t hb_rc_qcheck(const hb_hit_t *a, int32_t n_a, int32_t *f, int64_t *p, double bw_thres)
{ // lchain_qcheck, Hash_Table.cpp:1424-1473
	int32_t i, tot_g = 0, sc, dg, dq, dr, dd, span;
	if (n_a == 0) return -1;
	if (n_a > 1) {
		if (a[0].self_offset >= a[n_a - 1].self_offset || a[0].offset >= a[n_a - 1].offset) return -1;
		dq = (int32_t)a[n_a - 1].self_offset - (int32_t)a[0].self_offset; dr = (int32_t)a[n_a - 1].offset - (int32_t)a[0].offset;
		dd = dq >= dr ? dq - dr : dr - dq; dg = dq >= dr ? dr : dq;
		if (dg == 0 || (double)dd > (double)dg * bw_thres) return -1;
	}
	for (i = 1; i < n_a; ++i) { if (a[i - 1].self_offset >= a[i].self_offset) break; if (a[i - 1].offset >= a[i].offset) break; }
	if (i < n_a) return -1;
	const double bw_pen = 1.0 / bw_thres;
	f[0] = hb_rc_normal_w((int32_t)(a[0].cnt & 0xffu), (int32_t)(a[0].cnt >> 8)); p[0] = -1;
	for (i = 1; i < n_a; ++i) {
		dq = (int32_t)a[i].self_offset - (int32_t)a[i - 1].self_offset; dr = (int32_t)a[i].offset - (int32_t)a[i - 1].offset;
		dd = dq >= dr ? dq - dr : dr - dq; dg = dq >= dr ? dr : dq;
		if (dg == 0) break;
		tot_g += dd;
		if (dd > HB_THRE_MAX && (double)dd > (double)dg * bw_thres) break;
		span = (int32_t)(a[i].cnt & 0xffu);
		sc = dg < span ? dg : span;
		sc = hb_rc_normal_w(sc, (int32_t)(a[i].cnt >> 8));
#ifdef __CUDA_ARCH__
		sc -= (int32_t)__dmul_rn(__dmul_rn(__ddiv_rn((double)dd, (double)dg), bw_pen), (double)sc);
#else
		sc -= (int32_t)((((double)dd) / ((double)dg)) * bw_pen * ((double)sc));
#endif
		f[i] = f[i - 1] + sc; p[i] = i - 1;
	}
	if (i < n_a) return -1;
	if (n_a > 1) {
		dq = (int32_t)a[n_a - 1].self_offset - (int32_t)a[0].self_offset; dr = (int32_t)a[n_a - 1].offset - (int32_t)a[0].offset;
		dg = dq >= dr ? dr : dq; dd = tot_g;
		if ((double)dd > (double)dg * bw_thres) return -1;
	}
	return n_a;
}
HB_HD void hb_rc_rev_khit(hb_hit_t &an, int64_t xl, int64_t yl)
{ // rev_khit, Hash_Table.cpp:2287-2290 (uint32 arithmetic like the macro's assignment)
	an.self_offset = (uint32_t)(xl - 1 - ((int64_t)an.self_offset + 1 - (int64_t)(an.cnt & 0xffu)));
	an.offset = (uint32_t)(yl - 1 - ((int64_t)an.offset + 1 - (int64_t)(an.cnt & 0xffu)));
}
HB_HD void hb_rc_reverse_all(hb_hit_t *a, int64_t a_n, int64_t xl, int64_t yl)
{
	int64_t i; const int64_t h = a_n >> 1;
	for (i = 0; i < h; ++i) { const hb_hit_t z = a[i]; a[i] = a[a_n - i - 1]; a[a_n - i - 1] = z; hb_rc_rev_khit(a[i], xl, yl); hb_rc_rev_khit(a[a_n - i - 1], xl, yl); }
	if (a_n & 1) hb_rc_rev_khit(a[i], xl, yl);
}
// lchain_qdp_fix, Hash_Table.cpp:2293-2440 (quick_check = 1): the chain's indices land in t[0..cL)
HB_HD int64_t hb_rc_qdp_fix(hb_hit_t *a, int64_t a_n, int64_t *t, int64_t *p, int32_t *f, int64_t max_skip, int64_t max_iter, int64_t max_dis, double pen_gap, double pen_skip,
                            double bw_rate, int64_t xl, int64_t yl, int64_t left_fix, int64_t right_fix)
{
	int64_t max_f, n_skip, st, max_j, end_j, sc, msc = -1, msc_i = -1, max_ii, ovl, movl = INT32_MAX, i, j, cL = 0, must_p = 1, is_reorder = 0; int32_t mx, tmp;
	const int64_t ret = hb_rc_qcheck(a, (int32_t)a_n, f, p, bw_rate);
	if (ret > 0) { a_n = ret; msc_i = a_n - 1; msc = f[msc_i]; }
	else {
		for (j = 0; j < a_n; j++) t[j] = 0;
		if (right_fix && !left_fix) { hb_rc_reverse_all(a, a_n, xl, yl); is_reorder = 1; }
		if (!right_fix && !left_fix) must_p = 0;
		for (i = st = 0, max_ii = -1; i < a_n; ++i) {
			max_f = a[i].cnt & 0xffu; if (must_p) max_f = INT32_MIN;
			n_skip = 0; max_j = end_j = -1;
			if (i - st > max_iter) st = i - max_iter;
			for (j = i - 1; j >= 0; --j) {
				sc = hb_rc_sc_ff(a[i], a[j], bw_rate, pen_gap, pen_skip);
				if (sc == INT32_MIN) continue;
				sc += f[j];
				if (sc > max_f) { max_f = sc; max_j = j; if (n_skip > 0) --n_skip; }
				else if (t[j] == (int32_t)i) { if (++n_skip > max_skip) { if (max_j != -1 || must_p == 0) break; } }
				if (p[j] >= 0) t[p[j]] = i;
				if ((int64_t)a[i].self_offset > max_dis + (int64_t)a[j].self_offset) { if (max_j != -1) break; }
				if (j < st) { if (max_j != -1 || must_p == 0) break; }
			}
			end_j = j;
			if (max_ii < 0 || (int64_t)a[i].self_offset - (int64_t)a[max_ii].self_offset > max_dis) {
				mx = INT32_MIN; max_ii = -1;
				for (j = i - 1; j >= st && (int64_t)a[i].self_offset - (int64_t)a[j].self_offset <= max_dis; --j) if (mx < f[j]) { mx = f[j]; max_ii = j; }
			}
			if (max_ii >= 0 && max_ii < end_j) {
				tmp = hb_rc_sc_ff(a[i], a[max_ii], bw_rate, pen_gap, pen_skip);
				if (tmp != INT32_MIN && max_f < (int64_t)tmp + f[max_ii]) { max_f = (int64_t)tmp + f[max_ii]; max_j = max_ii; }
			}
			if (max_j == -1) { f[i] = 0; p[i] = max_j; } else { f[i] = (int32_t)max_f; p[i] = max_j; }
			if (max_ii < 0 || ((int64_t)a[i].self_offset - (int64_t)a[max_ii].self_offset <= max_dis && f[max_ii] < f[i])) max_ii = i;
			if (f[i] >= msc) {
				ovl = hb_chain_len(a[i].self_offset, a[i].self_offset, xl, a[i].offset, a[i].offset, yl);
				if (f[i] > msc || ovl < movl) { msc = f[i]; msc_i = i; movl = ovl; }
			}
		}
	}
	if (right_fix && left_fix) msc_i = a_n - 1;
	i = msc_i; cL = 0;
	while (i >= 0) { t[cL++] = i; msc_i = i; i = p[i]; }
	if (is_reorder) {
		hb_rc_reverse_all(a, a_n, xl, yl);
		for (i = 0; i < cL; i++) t[i] = a_n - t[i] - 1;
	} else for (i = 0, j = cL - 1; i < j; i++, j--) { const int64_t x = t[i]; t[i] = t[j]; t[j] = x; }
	return cL;
}
// lchain_simple, Hash_Table.cpp:2544-2604 with des == a: the unsigned arithmetic of `f[j] + a[i].cnt` (int32 + uint32) is the reference's
HB_HD int64_t hb_rc_lchain_simple(hb_hit_t *a, int64_t a_n, int64_t *t, int64_t *p, int32_t *f, int64_t max_skip, int64_t max_iter)
{
	if (a_n <= 0) return 0;
	int64_t max_f, n_skip, st, max_j, sc, msc = -1, msc_i = -1, i, j, cL = 0;
	for (i = 1, f[0] = (int32_t)a[0].cnt, p[0] = -1, msc_i = a_n - 1; i < a_n; i++) {
		j = i - 1;
		if (a[i].self_offset > a[j].self_offset && a[i].offset > a[j].offset) { p[i] = j; f[i] = (int32_t)((uint32_t)f[j] + a[i].cnt); }
		else break;
	}
	if (i < a_n) {
		for (j = 0; j < a_n; j++) t[j] = 0;
		f[0] = (int32_t)a[0].cnt; p[0] = -1; msc = f[0]; msc_i = 0;
		for (i = 1, st = 0; i < a_n; ++i) {
			max_f = INT32_MIN; n_skip = 0; max_j = -1;
			if (i - st > max_iter) st = i - max_iter;
			for (j = i - 1; j >= st; --j) {
				if (a[i].self_offset > a[j].self_offset && a[i].offset > a[j].offset) {
					sc = (int64_t)(uint32_t)((uint32_t)f[j] + a[i].cnt);
					if (sc > max_f) { max_f = sc; max_j = j; if (n_skip > 0) --n_skip; }
					else if (t[j] == (int32_t)i) { if (++n_skip > max_skip) break; }
					if (p[j] >= 0) t[p[j]] = i;
				}
			}
			f[i] = (int32_t)max_f; p[i] = max_j;
			if (f[i] > msc) { msc = f[i]; msc_i = i; }
		}
	}
	i = msc_i; cL = 0;
	while (i >= 0) { t[cL++] = i; i = p[i]; }
	for (i = 0, j = cL - 1; i < j; i++, j--) { const int64_t x = t[i]; t[i] = t[j]; t[j] = x; }
	for (i = 0; i < cL; i++) a[i] = a[t[i]]; // t ascending, t[i] >= i
	return cL;
}

// gen_single_khit, Correct.cpp:16865-17008: S.h[0..ch_n) = the chain; returns the number of points (0: nothing to align)
HB_HD int64_t hb_rc_single_khit(EcRc &S, int64_t ch_n, int64_t mode, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t max_skip, int64_t max_iter)
{
	hb_hit_t *ch_a = S.h; int64_t k, i, j, occ, m, ncn, srt = 1; const int64_t prefix = (mode == 0 || mode == 1) ? 1 : 0, suffix = (mode == 0 || mode == 2) ? 1 : 0;
	for (k = occ = m = 0; k < ch_n; k++) {
		if (!(ch_a[k].cnt & 0xffu)) continue;
		occ++; if ((ch_a[k].cnt & 0xffu) > 1) occ++;
		ch_a[m++] = ch_a[k];
	}
	ch_n = m; if (!ch_n) return 0;
	occ += prefix + suffix;
	ncn = occ;
	if (ncn > S.hcap) { S.ovf = 1; return 0; }
	hb_hit_t cht;
	if (suffix) { cht.self_offset = (uint32_t)qe; cht.offset = (uint32_t)te; cht.cnt = 1; cht.id_strand = 1; ch_a[--occ] = cht; }
	for (k = ch_n - 1; k >= 0; k--) { // expanded in place from the back, reading ch_a[k] between the writes exactly like the reference
		if (!(ch_a[k].cnt & 0xffu)) continue;
		cht.self_offset = ch_a[k].self_offset + 1; cht.offset = ch_a[k].offset + 1;
		cht.cnt = ch_a[k].cnt & 0xffu; cht.id_strand = cht.cnt;
		if ((int64_t)(ch_a[k].cnt & 0xffu) < (int64_t)S.h_khit) cht.id_strand = cht.cnt + 1;
		ch_a[--occ] = cht;
		if ((ch_a[k].cnt & 0xffu) > 1) {
			cht.self_offset = ch_a[k].self_offset + 1 - (ch_a[k].cnt & 0xffu); cht.offset = ch_a[k].offset + 1 - (ch_a[k].cnt & 0xffu);
			cht.cnt = ch_a[k].cnt & 0xffu; cht.id_strand = cht.cnt;
			if ((int64_t)(ch_a[k].cnt & 0xffu) < (int64_t)S.h_khit) cht.id_strand = cht.cnt + 1;
			ch_a[--occ] = cht;
		}
	}
	if (prefix) { cht.self_offset = (uint32_t)qs; cht.offset = (uint32_t)ts; cht.cnt = 1; cht.id_strand = 1; ch_a[--occ] = cht; }
	if (occ != 0) { hb_flag(S.err, 32); return 0; }
	ch_n = ncn;
	uint64_t q[2], t[2]; q[0] = q[1] = t[0] = t[1] = (uint64_t)-1;
	if (prefix) { q[0] = (uint64_t)qs; t[0] = (uint64_t)ts; }
	if (suffix) { q[1] = (uint64_t)qe; t[1] = (uint64_t)te; }
	for (k = m = occ = 0; k < ch_n; k++) { // the reference ASSIGNS the offset inside these two tests (`=`, not `==`)
		if (k > 0 && (uint64_t)ch_a[k].self_offset == q[0]) { ch_a[k].offset = (uint32_t)t[0]; if (ch_a[k].offset) continue; }
		if (k + 1 < ch_n && (uint64_t)ch_a[k].self_offset == q[1]) { ch_a[k].offset = (uint32_t)t[1]; if (ch_a[k].offset) continue; }
		if (m > 0) {
			if (ch_a[k].self_offset > ch_a[m - 1].self_offset && ch_a[k].offset > ch_a[m - 1].offset) occ++;
			if (ch_a[k].self_offset <= ch_a[m - 1].self_offset) srt = 0;
		} else occ++;
		ch_a[m++] = ch_a[k];
	}
	ch_n = m;
	if (occ == ch_n) return ch_n;
	if (!srt) {
		if (hb_rs_sort32(ch_a, ch_a + ch_n, [](const hb_hit_t &x) -> uint32_t { return x.self_offset; }, S.rs)) { S.ovf = 1; return 0; }
		for (i = 1, j = 0; i <= ch_n; i++) {
			if (i == ch_n || ch_a[i].self_offset != ch_a[j].self_offset) {
				if (i - j > 1) { if (hb_rs_sort32(ch_a + j, ch_a + i, [](const hb_hit_t &x) -> uint32_t { return x.offset; }, S.rs)) { S.ovf = 1; return 0; } }
				j = i;
			}
		}
	}
	occ = ch_n;
	ch_n = hb_rc_lchain_simple(ch_a + prefix, ch_n - prefix - suffix, S.t, S.p, S.f, max_skip, max_iter);
	ch_n += prefix + suffix; if (suffix) ch_a[ch_n - 1] = ch_a[occ - 1];
	return ch_n;
}

// gen_win_chain, Correct.cpp:17012-17075 (is_accurate = 1): z = the PRIVATE copy of the overlap with step A's windows
HB_HD int64_t hb_rc_win_chain(EcBCtx &C, EcZ &z, EcRc &S, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode)
{
	const int64_t wl = C.w_l, wn = z.wn; int64_t k, ws, we, os, oe, wsk, occ = 0, ch_n; uint32_t w = 1;
	S.hn = 0;
	ws = qs; if (ws < z.x_pos_s) ws = z.x_pos_s;
	wsk = ((ws / wl) * wl - (z.x_pos_s / wl) * wl) / wl; // get_win_id_by_s, Correct.h:1306
	for (wsk = (wsk < wn ? wsk : wn - 1); wsk < wn && wsk >= 0 && qs > z.w[wsk].x_end; wsk++);
	for (wsk = (wsk < wn ? wsk : wn - 1); wsk >= 0 && qs < z.w[wsk].x_start; wsk--);
	if (wsk < 0) wsk = 0;
	if (mode == 0 || mode == 1) hb_rc_push_khit(S, (int32_t)qs, (int32_t)ts, 0, &w);
	EcCtx A; A.R = S.R; A.q = C.q; A.t = C.t; A.e_rate = C.e_rate; A.w_l = C.w_l; A.ez.path = S.path1; A.ez.cig = S.cig1; A.ez.cn = 0;
	A.pool = S.zc; A.pool_used = S.zc_used; A.pool_cap = S.zc_cap; A.err = S.err;
	for (k = wsk; k < wn && z.w[k].x_start < qe; k++) {
		if (z.w[k].y_end == -1) continue;
		ws = z.w[k].x_start; we = (int64_t)z.w[k].x_end + 1;
		os = qs > ws ? qs : ws; oe = qe < we ? qe : we;
		if (!(oe > os)) continue;
		if (!z.w[k].clen) {
			if (hb_gen_backtrace_adv(A, z, &z.w[k]) && z.w[k].clen) z.w[k].cidx |= HB_RC_PRIV;
			if (*S.zc_used > S.zc_cap) { S.ovf = 1; return 0; }
		}
		const uint16_t *cg = (z.w[k].cidx & HB_RC_PRIV) ? S.zc + (z.w[k].cidx & ~HB_RC_PRIV) : S.poolA + z.w[k].cidx;
		occ += hb_rc_extract(S, cg, z.w[k].clen, z.w[k].y_start, z.w[k].x_start, (int32_t)ts, (int32_t)te, (int32_t)qs, (int32_t)qe, 10, wl);
	}
	if (mode == 0 || mode == 2) hb_rc_push_khit(S, (int32_t)(qe - 1), (int32_t)(te - 1), 0, &w);
	if (S.ovf) return 0;
	if (!occ) { S.hn = 0; return 0; }
	const int64_t ch_n0 = S.hn;
	ch_n = hb_rc_qdp_fix(S.h, ch_n0, S.t, S.p, S.f, 25, 5000, HB_MAX_SIN_L >> 1, S.pen_gap, S.pen_skip, C.e_rate, C.ql, C.tl, (mode == 0 || mode == 1) ? 1 : 0, (mode == 0 || mode == 2) ? 1 : 0);
	for (k = occ = 0; k < ch_n; k++) { S.h[k] = S.h[S.t[k]]; if (S.h[k].cnt & 0xffu) occ++; }
	if (occ <= 0) return 0;
	return hb_rc_single_khit(S, ch_n, mode, qs, qe, ts, te, 25, 5000);
}

HB_HD bool hb_rc_is_ualn(const hb_wl_t &u) { return u.error == INT16_MAX && u.clen == 0 && u.extra_end < 0; } // is_ualn_win, Correct.h
// rechain_aln_hc, Correct.cpp:17669-17748, for the unaligned window aux_i of the list under construction
HB_HD void hb_rc_window(EcBCtx &C, EcZ &z, EcRc &S, int32_t aux_i)
{
	const int64_t qs = C.aw[aux_i].x_start, qe = (int64_t)C.aw[aux_i].x_end + 1, ts = C.aw[aux_i].y_start, te = (int64_t)C.aw[aux_i].y_end + 1;
	if (qe - qs < HB_FORCE_SIN_L || te - ts < HB_FORCE_SIN_L) return;
	const int64_t mode = C.aw[aux_i].error_threshold;
	if (mode < 0 || mode > 2) { hb_flag(S.err, 32); return; }
	const int64_t ch_n = hb_rc_win_chain(C, z, S, qs, qe, ts, te, mode);
	if (S.ovf || !ch_n) return;
	const hb_hit_t *ch_a = S.h; int todo = 1;
	if (mode == 0) { if (ch_n <= 2) todo = 0; }
	else if (ch_n <= 1) todo = 0;
	if (!todo) return;
	const int32_t an0 = C.awn;
	// hc_ovlp_base_direct with pre_mode = mode (Correct.cpp:17461-17514): the pieces between the points, not beyond the fixed ends
	int64_t si = 0, ei = ch_n, uq[2], ut[2], um;
	if (mode == 0) { si = 1; ei = ch_n - 1; } else if (mode == 1) si = 1; else ei = ch_n - 1;
	for (int64_t i = si; i <= ei && !C.ez.ovf; i++) {
		const int st = hb_ecb_segment(C, z, ch_a, ch_n, i, uq, ut, &um);
		if (C.ez.ovf) break;
		hb_ecb_apply(C, st, hb_aln_of(C.ez), uq, ut, um);
	}
	if (C.ez.ovf) return;
	const int32_t an = C.awn;
	if (an == an0 + 1 && !hb_rc_is_ualn(C.aw[an - 1])) { // one aligned window over the whole old one replaces it
		const hb_wl_t &o = C.aw[aux_i], &n = C.aw[an - 1];
		const int q0 = o.x_start == n.x_start, q1 = o.x_end == n.x_end, t0 = o.y_start == n.y_start, t1 = o.y_end == n.y_end;
		if ((mode == 0 && q0 && q1 && t0 && t1) || (mode == 1 && q0 && t0) || (mode == 2 && q1 && t1)) {
			hb_b_flush(C); // its cigar leaves the per-thread buffer before the record moves
			C.aw[aux_i] = C.aw[an - 1]; C.awn--;
		}
	}
}

// The rescue of one accepted overlap after its windows were built (hb_ecb_finish ran: out / C.aw hold the result, every cigar flushed).
// zA = the overlap with step A's list.  On success out is the reference's state after gen_hc_fast_cigar0 (+ reassign_gaps when
// C.do_gaps) and out->need_rechain = 0; when a scratch is too small nothing of the overlap is final: need_rechain stays 1.
HB_HD void hb_ecb_rechain(EcBCtx &C, const EcZ &zA, int64_t re_A, EcRc &S, hb_alnb_t *out)
{
	if (out->st != 2 || !out->need_rechain) return;
	if (zA.wn > S.zcap) return;
	EcZ z = zA; z.w = S.zw;
	for (int32_t k = 0; k < zA.wn; k++) S.zw[k] = zA.w[k];
	*S.zc_used = 0; S.ovf = 0; S.hn = 0;
	C.awn = (int32_t)out->w_n; C.open = -1; C.wcn = 0; C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0; C.re_A = re_A; C.no_myers = 0;
	C.gap_re = re_A - out->nh_err; // what reassign_gaps removed so far (0 when step C is not fused)
	const int32_t aux_n = C.awn;
	for (int32_t i = 0; i < aux_n && !S.ovf && !C.ez.ovf; i++) if (hb_rc_is_ualn(C.aw[i])) hb_rc_window(C, z, S, i);
	hb_b_flush(C);
	if (S.ovf || C.ez.ovf) { out->need_rechain = 1; return; } // reported: the read's status says its lists are not final
	if (C.awn > aux_n) {
		int32_t m = 0;
		for (int32_t i = 0; i < C.awn; i++) { if (i < aux_n && hb_rc_is_ualn(C.aw[i])) continue; C.aw[m++] = C.aw[i]; }
		C.awn = m;
		for (int32_t i = 1; i < m; i++) { // radix_sort_window_list_xs_srt: x_start is unique, any sort gives the same list
			const hb_wl_t v = C.aw[i]; int32_t j = i;
			for (; j > 0 && C.aw[j - 1].x_start > v.x_start; j--) C.aw[j] = C.aw[j - 1];
			C.aw[j] = v;
		}
	}
	hb_ecb_finish(C, zA, re_A, out);
	out->need_rechain = 0;
}
