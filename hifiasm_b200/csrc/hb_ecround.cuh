// hb_ecround.cuh — the per-read steps that close an error-correction round.
//
// cal_ec_r (ecovlp.cpp:6268) after the alignment / phasing / consensus of a read:
//   a15  push_ne_ovlp (2585) + extract_max_exact (2520, _sub 2437) + check_well_cal (2750): the round's paf[i] with the longest
//        exact interval of every same-haplotype overlap mapped through the read's edit script, is_fully_corrected / is_abnormal
//   a16  worker_sl_ec (5965): apply the edit script -> new 2-bit read, new length, new N list
//   a17  worker_update_dc_ec (3808) -> quick_exact_match (3521) / adjust_exact_match (3454): remap each exact interval through the
//        TARGET's edit script, extend to the read ends, compare the packed substrings
//   a18  worker_hap_post_rev (3866) / flip_paf_rc (3845): reverse-complement the read, flip both overlap lists
// Edit scripts are the reference's own encoding (push_trace_bp_f, Levenshtein_distance.h:640): uint16 runs
//   op = w >> 14: 0 match (len 14 bits), 1 mismatch (target base 2 bits, query base 2 bits, len 10 bits),
//   2 insertion into the read (base 2 bits, len 12 bits), 3 deletion from the read (base 2 bits, len 12 bits).
// All bodies are one-thread-per-unit (HB_HD): integer work on the packed reads in HBM.
#pragma once
#include "hb_common.cuh"
#include "hb_ecaln.cuh"
#include "hb_ecphase.cuh"

struct ScRun { uint32_t op, bq, bt, len; };
HB_HD bool hb_ualn_w(const hb_wl_t &u) { return u.error == INT16_MAX && u.clen == 0 && u.extra_end < 0; } // is_ualn_win, Correct.h:1362
#define HB_SC_NONE 0xffffu
HB_HD void hb_sc_entry(uint16_t w, uint32_t *c, uint32_t *bq, uint32_t *bt, uint32_t *len)
{
	*c = w >> 14; *bq = *bt = HB_SC_NONE;
	if (*c == 2 || *c == 3) { *bt = (w >> 12) & 3; *len = w & 0xfff; }
	else if (*c == 1) { *bt = (w >> 12) & 3; *bq = (w >> 10) & 3; *len = w & 0x3ff; }
	else *len = w & 0x3fff;
}
// pop_trace_bp_f, Levenshtein_distance.h:672: one run = consecutive entries of the same op and bases
HB_HD uint32_t hb_sc_pop(const uint16_t *a, uint32_t n, uint32_t i, ScRun *r)
{
	uint32_t c, bq, bt, len, sc, sbq, sbt, sl;
	hb_sc_entry(a[i], &c, &bq, &bt, &len);
	for (i++; i < n && c == (uint32_t)(a[i] >> 14); i++) {
		hb_sc_entry(a[i], &sc, &sbq, &sbt, &sl);
		if (bq != sbq || bt != sbt) break;
		len += sl;
	}
	if (c == 3) { bq = bt; bt = HB_SC_NONE; }
	r->op = c; r->bq = bq; r->bt = bt; r->len = len;
	return i;
}
// pop_trace_bp_rev_f, Levenshtein_distance.h:709
HB_HD int64_t hb_sc_pop_rev(const uint16_t *a, int64_t i, ScRun *r)
{
	uint32_t c, bq, bt, len, sc, sbq, sbt, sl;
	hb_sc_entry(a[i], &c, &bq, &bt, &len);
	for (i--; i >= 0 && c == (uint32_t)(a[i] >> 14); i--) {
		hb_sc_entry(a[i], &sc, &sbq, &sbt, &sl);
		if (bq != sbq || bt != sbt) break;
		len += sl;
	}
	if (c == 3) { bq = bt; bt = HB_SC_NONE; }
	r->op = c; r->bq = bq; r->bt = bt; r->len = len;
	return i;
}
// pop_trace, Levenshtein_distance.h:533 (alignment cigars: op << 14 | len, runs of the same op merged)
HB_HD uint32_t hb_cg_pop(const uint16_t *a, uint32_t n, uint32_t i, uint32_t *c, uint32_t *len)
{
	*c = a[i] >> 14; *len = a[i] & 0x3fff;
	for (i++; i < n && *c == (uint32_t)(a[i] >> 14); i++) *len += a[i] & 0x3fff;
	return i;
}

// ---- a16: worker_sl_ec (ecovlp.cpp:5965-6061) ------------------------------------------------------------------------------------------
// first walk (5973-5980): length of the corrected read; a script without any edit leaves the read untouched
HB_HD void hb_sl_len(const uint16_t *sc, uint32_t n, uint32_t old_len, uint32_t *new_len, uint32_t *changed)
{
	uint32_t i = 0, yk = 0, tot_e = 0; ScRun r;
	while (i < n) { i = hb_sc_pop(sc, n, i, &r); if (r.op != 3) yk += r.len; if (r.op != 0) tot_e += r.len; }
	*changed = tot_e ? 1 : 0; *new_len = tot_e ? yk : old_len;
}
// second walk (5990-6012) + ha_compress_base (Process_Read.cpp:792): dst = the read's packed bytes (len/4+1, pad bits zero),
// dst_npos = its N positions (at most as many as the source read has: edits only write A/C/G/T); returns the number of Ns
HB_HD uint32_t hb_sl_apply(const RdView &src, const uint16_t *sc, uint32_t n, uint8_t *dst, uint32_t *dst_npos)
{
	uint32_t i = 0, xk = 0, yk = 0, nn = 0, acc = 0; ScRun r;
	while (i < n) {
		i = hb_sc_pop(sc, n, i, &r);
		if (r.op == 0 || r.op == 1 || r.op == 2) {
			for (uint32_t k = 0; k < r.len; k++) {
				int b = r.op == 0 ? src.at((int64_t)xk + k) : (int)r.bt;
				if (b == 4) { dst_npos[nn++] = yk; b = 0; }
				acc |= (uint32_t)b << ((3 - (yk & 3)) << 1);
				if ((yk & 3) == 3) { dst[yk >> 2] = (uint8_t)acc; acc = 0; }
				yk++;
			}
		}
		if (r.op != 2) xk += r.len;
	}
	dst[yk >> 2] = (uint8_t)acc; // the partial byte, or the extra byte of a length divisible by four (left undefined by the reference, zero here)
	return nn;
}

// ---- a18: worker_hap_post_rev (ecovlp.cpp:3866-3908) -----------------------------------------------------------------------------------
// byte j of the reverse complement of a packed read (N positions hold code 0 on both strands)
HB_HD uint8_t hb_rc_byte(const RdView &src, uint32_t j)
{
	uint32_t acc = 0; const uint32_t L = src.len;
	for (uint32_t t = 0; t < 4; t++) {
		const uint32_t p = 4 * j + t; if (p >= L) break;
		const uint32_t fp = L - 1 - p; int b = 3 - hb_base(src.p, fp);
		if (src.nn) { uint32_t lo = 0, hi = src.nn; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (src.npos[mid] < fp) lo = mid + 1; else hi = mid; } if (lo < src.nn && src.npos[lo] == fp) b = 0; }
		acc |= (uint32_t)b << ((3 - t) << 1);
	}
	return (uint8_t)acc;
}
// flip_paf_rc (ecovlp.cpp:3845-3864): in place, returns the new length of the list
HB_HD uint32_t hb_flip_paf(const uint32_t *rlen, uint64_t rid, hb_ma_hit_t *a, uint32_t n)
{
	uint32_t m = 0; const int64_t ql = rlen[rid];
	for (uint32_t k = 0; k < n; k++) {
		hb_ma_hit_t z = a[k]; const int64_t tl = rlen[z.tn];
		int64_t qs = (uint32_t)z.qns, qe = z.qe, ts = z.ts, te = z.te;
		if (qs > ql) qs = ql; if (qe > ql) qe = ql; if (ts > tl) ts = tl; if (te > tl) te = tl;
		if (qe > qs && te > ts) {
			z.qns = (z.qns >> 32 << 32) | (uint64_t)(ql - qe); z.qe = (uint32_t)(ql - qs); z.ts = (uint32_t)(tl - te); z.te = (uint32_t)(tl - ts);
			a[m++] = z;
		}
	}
	return m;
}

// ---- a17: worker_update_dc_ec (ecovlp.cpp:3808-3842) -----------------------------------------------------------------------------------
// adjust_exact_match (3454-3519): x = coordinates on the target before its correction, y = after; the longest piece of [xs0, xe0)
// inside one match run of the target's script, mapped to y; rev walks the script from its end (the interval is on the reverse strand)
HB_HD uint32_t hb_adjust_exact_match(const uint16_t *sc, int64_t cn, int64_t xs0, int64_t xe0, int64_t ys0, int64_t ye0, uint64_t *rxs, uint64_t *rxe, uint64_t *rys, uint64_t *rye, uint32_t rev)
{
	*rxs = *rxe = *rys = *rye = 0;
	if (xe0 <= xs0 || ye0 <= ys0) return 0;
	int64_t xk = 0, yk = 0, ck, wx0, wy0, wx1, os, oe; ScRun r; uint64_t ovlp;
	if (!rev) {
		ck = 0;
		while (ck < cn && xk < xe0) {
			wx0 = xk; wy0 = yk;
			ck = hb_sc_pop(sc, (uint32_t)cn, (uint32_t)ck, &r);
			if (r.op != 2) xk += r.len; if (r.op != 3) yk += r.len;
			wx1 = xk;
			if (r.op == 0) {
				os = xs0 > wx0 ? xs0 : wx0; oe = xe0 < wx1 ? xe0 : wx1; ovlp = oe > os ? (uint64_t)(oe - os) : 0;
				if (ovlp > 0 && ovlp > *rxe - *rxs) { *rxs = (uint64_t)(wy0 + os - wx0); *rxe = (uint64_t)(wy0 + oe - wx0); *rys = (uint64_t)(ys0 + os - xs0); *rye = (uint64_t)(ys0 + oe - xs0); }
			}
		}
	} else {
		ck = cn - 1;
		while (ck >= 0 && xk < xe0) {
			wx0 = xk; wy0 = yk;
			ck = hb_sc_pop_rev(sc, ck, &r);
			if (r.op != 2) xk += r.len; if (r.op != 3) yk += r.len;
			wx1 = xk;
			if (r.op == 0) {
				os = xs0 > wx0 ? xs0 : wx0; oe = xe0 < wx1 ? xe0 : wx1; ovlp = oe > os ? (uint64_t)(oe - os) : 0;
				if (ovlp > 0 && ovlp > *rxe - *rxs) { *rxs = (uint64_t)(wy0 + os - wx0); *rxe = (uint64_t)(wy0 + oe - wx0); *rys = (uint64_t)(ys0 + os - xs0); *rye = (uint64_t)(ys0 + oe - xs0); }
			}
		}
	}
	return (uint32_t)(*rxe - *rxs);
}
// quick_exact_match (3521-3586) on the corrected read store R; sc / sc_off = all edit scripts of the round (scc)
HB_HD uint32_t hb_quick_exact_match(const DevReads &R, uint64_t qid, hb_ma_hit_t *z, const uint16_t *sc, const uint64_t *sc_off)
{
	uint64_t rts, rte, rqs, rqe, f = 0; int64_t ql, tl, qr, tr, qs, qe, ts, te;
	if (hb_adjust_exact_match(sc + sc_off[z->tn], (int64_t)(sc_off[z->tn + 1] - sc_off[z->tn]), z->ts, z->te, (uint32_t)z->qns, z->qe, &rts, &rte, &rqs, &rqe, z->rev)) {
		z->ts = (uint32_t)rts; z->te = (uint32_t)rte; f = 1;
		z->qns = (z->qns >> 32 << 32) | (uint64_t)(uint32_t)rqs; z->qe = (uint32_t)rqe;
	}
	ql = R.len[qid]; tl = R.len[z->tn];
	qs = (uint32_t)z->qns; qe = z->qe; ts = z->ts; te = z->te;
	if (qs >= ql) qs = ql; if (qe > ql) qe = ql; if (qe <= qs) f = 0;
	if (ts >= tl) ts = tl; if (te > tl) te = tl; if (te <= ts) f = 0;
	if ((qe - qs) != (te - ts)) f = 0;
	if (qs <= ts) { ts -= qs; qs = 0; } else { qs -= ts; ts = 0; }
	qr = ql - qe; tr = tl - te;
	if (qr <= tr) { qe = ql; te += qr; } else { te = tl; qe += tr; }
	z->qns = (z->qns >> 32 << 32) | (uint64_t)(uint32_t)qs; z->qe = (uint32_t)qe; z->ts = (uint32_t)ts; z->te = (uint32_t)te;
	if (f && (te - ts) == (qe - qs) && qe > qs) {
		const RdView Q = hb_rd_view(R, qid, 0), T = hb_rd_view(R, z->tn, z->rev);
		if (hb_seq_equal(Q, qs, T, ts, qe - qs)) return 1;
	}
	return 0;
}
// one record of paf[i]; returns 1 when the record stays exact
HB_HD uint32_t hb_update_dc(const DevReads &R, uint64_t qid, hb_ma_hit_t *z, const uint16_t *sc, const uint64_t *sc_off)
{
	if (z->el && hb_quick_exact_match(R, qid, z, sc, sc_off)) { z->el = 1; return 1; }
	z->el = 0; return 0;
}

// ---- a15: push_ne_ovlp(flag 1, ec) (ecovlp.cpp:2585-2638) -------------------------------------------------------------------------------
// extract_max_exact_sub (2437-2484): ec = the read's edit script (x = before, y = after the correction); the cursor (xk, yk, ck)
// persists across calls of one overlap.  [xs0, xe0) = a match run of the overlap on the uncorrected read, ys0 = its target start.
struct EcCur { int64_t xk, yk, ck; };
HB_HD uint32_t hb_max_exact_sub(const uint16_t *ec, int64_t cn, int64_t xs0, int64_t xe0, int64_t ys0, EcCur *cur, uint64_t *rxs, uint64_t *rxe, uint64_t *rys, uint64_t *rye)
{
	int64_t xk = cur->xk, yk = cur->yk, ck = cur->ck, ol, wx0, wy0, wx1, os, oe; uint32_t op; ScRun r; uint32_t ovlp;
	*rxs = *rxe = *rys = *rye = 0;
	if (ck < 0 || ck > cn) { ck = 0; xk = 0; yk = 0; }
	while (ck > 0 && xk >= xs0) { // back up entry by entry (not run by run)
		--ck; op = ec[ck] >> 14;
		if (op == 2 || op == 3) ol = ec[ck] & 0xfff; else if (op == 1) ol = ec[ck] & 0x3ff; else ol = ec[ck] & 0x3fff;
		if (op != 2) xk -= ol; if (op != 3) yk -= ol;
	}
	while (ck < cn && xk < xe0) {
		wx0 = xk; wy0 = yk;
		ck = hb_sc_pop(ec, (uint32_t)cn, (uint32_t)ck, &r);
		if (r.op != 2) xk += r.len; if (r.op != 3) yk += r.len;
		wx1 = xk;
		if (r.op == 0) {
			os = xs0 > wx0 ? xs0 : wx0; oe = xe0 < wx1 ? xe0 : wx1; ovlp = oe > os ? (uint32_t)(oe - os) : 0;
			if (ovlp > 0 && ovlp > *rxe - *rxs) { *rxs = (uint64_t)(wy0 + os - wx0); *rxe = (uint64_t)(wy0 + oe - wx0); *rys = (uint64_t)(ys0 + os - xs0); *rye = (uint64_t)(ys0 + oe - xs0); }
		}
	}
	cur->xk = xk; cur->yk = yk; cur->ck = ck;
	return *rxe > *rxs ? 1 : 0;
}
// extract_max_exact (2520-2583): w[0..wn) = the overlap's windows after reassign_gaps, pool = their cigars
HB_HD uint32_t hb_max_exact(const hb_wl_t *w, uint32_t wn, const uint16_t *pool, const uint16_t *ec, int64_t ecn, uint32_t *rxs, uint32_t *rxe, uint32_t *rys, uint32_t *rye)
{
	*rxs = *rxe = *rys = *rye = 0xffffffffu;
	uint64_t rx0, rx1, ry0, ry1, xk, yk, wx0, wy0, wx1, mx0 = 0, mx1 = 0, my0 = 0, my1 = 0; uint32_t cl, ck, c; EcCur cur; cur.xk = cur.yk = cur.ck = 0;
	for (uint32_t k = 0; k < wn; k++) {
		const uint16_t *ct = pool + w[k].cidx; const uint32_t cn = w[k].clen; ck = 0;
		xk = (uint64_t)(int64_t)w[k].x_start; yk = (uint64_t)(int64_t)w[k].y_start;
		while (ck < cn) {
			wx0 = xk; wy0 = yk;
			ck = hb_cg_pop(ct, cn, ck, &c, &cl);
			if (c != 2) xk += cl; if (c != 3) yk += cl;
			wx1 = xk;
			if (wx1 <= wx0) continue;
			if ((wx1 - wx0) <= (mx1 - mx0)) continue;
			if (c == 0 && hb_max_exact_sub(ec, ecn, (int64_t)wx0, (int64_t)wx1, (int64_t)wy0, &cur, &rx0, &rx1, &ry0, &ry1)) {
				if ((rx1 - rx0) > (mx1 - mx0)) { mx0 = rx0; mx1 = rx1; my0 = ry0; my1 = ry1; }
			}
		}
	}
	if (mx1 > mx0) { *rxs = (uint32_t)mx0; *rxe = (uint32_t)mx1; *rys = (uint32_t)my0; *rye = (uint32_t)my1; return 1; }
	return 0;
}
// the large-indel test wcns_gen makes per same-haplotype overlap (ecovlp.cpp:2299-2360): 1 when no gap between / around the windows and
// no indel run inside them reaches 6 bases (ma_hit_t.no_l_indel)
HB_HD uint32_t hb_without_large_indel(const hb_wl_t *w, uint32_t wn, const uint16_t *pool, int64_t x_pos_s, int64_t x_pos_e)
{
	uint32_t l_nid = 0; int64_t li = -1; uint64_t p0, p1, n_id;
	if (!wn) return 0;
	for (uint32_t i = 0; i < wn; i++) {
		if (hb_ualn_w(w[i])) { n_id = (uint64_t)((int64_t)w[i].x_end + 1 - w[i].x_start); if (n_id >= 6) l_nid = 1; continue; }
		if (l_nid == 0) {
			if (i == 0) { p0 = (uint64_t)(int64_t)w[i].x_start; p1 = (uint64_t)x_pos_s; n_id = p0 >= p1 ? p0 - p1 : p1 - p0; if (n_id >= 6) l_nid = 1; }
			if (li != -1) {
				p0 = (uint64_t)(int64_t)w[i].x_start; p1 = (uint64_t)((int64_t)w[li].x_end + 1); n_id = p0 >= p1 ? p0 - p1 : p1 - p0; if (n_id >= 6) l_nid = 1;
				p0 = (uint64_t)(int64_t)w[i].y_start; p1 = (uint64_t)((int64_t)w[li].y_end + 1); n_id = p0 >= p1 ? p0 - p1 : p1 - p0; if (n_id >= 6) l_nid = 1;
			}
			if (i + 1 == wn) { p0 = (uint64_t)(int64_t)w[i].x_end; p1 = (uint64_t)x_pos_e; n_id = p0 >= p1 ? p0 - p1 : p1 - p0; if (n_id >= 6) l_nid = 1; }
			if (l_nid == 0) {
				const uint16_t *ct = pool + w[i].cidx; const uint32_t cn = w[i].clen; uint32_t ci = 0, c, cl;
				while (ci < cn && l_nid == 0) { ci = hb_cg_pop(ct, cn, ci, &c, &cl); if (c >= 2 && cl >= 6) l_nid = 1; }
			}
		}
		li = (int64_t)i;
	}
	return l_nid ? 0 : 1;
}
// check_well_cal (ecovlp.cpp:2750-2801): coverage sweep over the emitted records' query intervals; srt = scratch of 2 n words
HB_HD void hb_check_well_cal(const uint16_t *sc, uint32_t scn, uint64_t *srt, const hb_ma_hit_t *a, uint32_t n, int64_t len, int64_t min_dp, uint8_t *f_ec, uint8_t *abnormal)
{
	uint32_t k, m = 0; int64_t dp = 0, old_dp, st = 0, ed = 0;
	*f_ec = 1; *abnormal = 0;
	for (k = 0; k < n; k++) { const uint64_t s = (uint32_t)a[k].qns, e = a[k].qe; srt[m++] = s << 1; srt[m++] = (e << 1) | 1; }
	hb_heapsort64(srt, m); // radix_sort_ec64 on bare keys: any sort gives the same array
	for (k = 0; k < m; k++) {
		old_dp = dp;
		if (srt[k] & 1) --dp; else ++dp;
		ed = (int64_t)(srt[k] >> 1);
		if (ed > st) {
			if (old_dp < min_dp) *f_ec = 0;
			if (old_dp == 0) { if (st > 0 && ed < len) *abnormal = 1; else if (*abnormal == 0) *abnormal = 2; }
		}
		st = ed;
	}
	ed = len; old_dp = dp;
	if (ed > st) {
		if (old_dp < min_dp) *f_ec = 0;
		if (old_dp == 0) { if (st > 0 && ed < len) *abnormal = 1; else if (*abnormal == 0) *abnormal = 2; }
	}
	if (*f_ec) { for (k = 0; k < scn && (sc[k] >> 14) == 0; k++) {} if (k < scn) *f_ec = 0; }
}

// the round's paf[i]: the same-haplotype overlaps (is_match == 1) of the de-duplicated list, in list order.  ph / alnb = the read's overlaps
// (hb_chains order) after phasing / after step C; ord[0..keep) from hb_ec_dedup; wl / pool = window lists and cigars of step C; ec = the read's
// edit script (ecn entries).  out = up to keep records.
HB_HD uint32_t hb_ec_source_list(const DevReads &R, uint64_t qid, const hb_phase_t *ph, const hb_alnb_t *alnb, const PhPair *ord, uint32_t keep,
                                 const hb_wl_t *wl, const uint16_t *pool, const uint16_t *ec, int64_t ecn, hb_ma_hit_t *out)
{
	uint32_t no = 0;
	for (uint32_t k = 0; k < keep; k++) {
		const hb_phase_t &z = ph[ord[k].idx]; const hb_alnb_t &b = alnb[ord[k].idx];
		if (z.is_match != 1) continue;
		const hb_wl_t *w = wl + b.w_off; uint32_t rxs, rxe, rys, rye;
		hb_ma_hit_t h; h.qns = (qid << 32) | z.x_pos_s; h.qe = z.x_pos_e + 1; h.tn = z.y_id; h.ts = z.y_pos_s; h.te = z.y_pos_e + 1; h.rev = z.rev;
		h.bl = R.len[z.y_id]; h.ml = (uint32_t)z.strong; h.no_l_indel = (uint8_t)hb_without_large_indel(w, b.w_n, pool, z.x_pos_s, z.x_pos_e); h.del = 0; h.el = 0;
		for (int t = 0; t < 6; t++) h.pad[t] = 0;
		hb_max_exact(w, b.w_n, pool, ec, ecn, &rxs, &rxe, &rys, &rye);
		if (rxe > rxs) { h.qns = (qid << 32) | rxs; h.qe = rxe; h.ts = rys; h.te = rye; h.el = 1; }
		out[no++] = h;
	}
	return no;
}
