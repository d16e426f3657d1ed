// hb_ecaln.cuh — alignment stage of an error-correction round, per overlap.
//
// What gen_hc_r_alin (Correct.cpp:25617-25675) does to an overlap before the base-level CIGAR:
//   align_hc_ed_post_extz (12951)  window pass + gap filling (push_hc_wlst_exz 12776: forward from the previous
//                                  window's end, backward from this window's traced start) + the 0.9 aligned-fraction cut
//   gen_extend_err_exz   (13400)   error estimate over the still unaligned windows by extension from their neighbours
//                                  (gen_extend_err_0_exz 13224, ed_cut 13063 -> Reserve_Banded_BPM_Extension[_REV])
// The initial per-window alignments come from k_windows (one thread per window); this file is the per-overlap
// state machine that consumes them (one thread per overlap: the steps depend on each other), reading bases
// straight from the 2-bit packed reads.  Everything is integer work except three double comparisons.
#pragma once
#include "hb_common.cuh"

#define HB_THRE_MAX 31        // THRESHOLD_MAX_SIZE, Hash_Table.h:24
#define HB_OVLP_CUT 0.9       // OVERLAP_THRESHOLD_HIFI_FILTER, Hash_Table.h:19
#define HB_EC_CIG_TMP 160     // u16 entries of per-thread cigar scratch (a window cigar has <= 2*31+2 runs)

// base of a read on a strand: 0..3, 4 = N
struct RdView {
	const uint8_t *p; const uint32_t *npos; uint32_t nn, len, rev;
	HB_HD int at(int64_t j) const
	{
		const uint32_t fp = rev ? (uint32_t)(len - 1 - j) : (uint32_t)j;
		if (nn) { uint32_t lo = 0, hi = nn; while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (npos[mid] < fp) lo = mid + 1; else hi = mid; } if (lo < nn && npos[lo] == fp) return 4; }
		const int b = hb_base(p, fp); return rev ? 3 - b : b;
	}
};
HB_HD RdView hb_rd_view(const DevReads &R, uint64_t id, uint32_t rev)
{
	RdView v; v.p = R.packed + R.off[id]; v.npos = R.npos + R.noff[id]; v.nn = (uint32_t)(R.noff[id + 1] - R.noff[id]); v.len = R.len[id]; v.rev = rev;
	return v;
}

// bit_extz_t (Levenshtein_distance.h:776-785): the fields this path uses; cigar / path live in per-thread scratch
struct EcEz { int32_t ps, pe, pl, ts, te, tl, thre, err; uint16_t *cig; int32_t cn; uint64_t *path; };

// per-overlap working state: overlap_region + window_list_alloc (Hash_Table.h:64-106)
struct EcZ {
	int64_t x_pos_s, x_pos_e, y_pos_s; uint32_t y_id, rev;
	const uint64_t *fc; uint32_t fc_n;
	int64_t align_length;
	hb_wl_t *w; int32_t wn;       // window list (capacity = number of windows of the overlap)
};
struct EcCtx {
	DevReads R; RdView q, t;      // query read (forward), target read (on the overlap's strand)
	double e_rate; int64_t w_l;
	EcEz ez;
	uint16_t *pool; unsigned long long *pool_used; uint64_t pool_cap; int *err; // cigar pool (bump allocation), error flags
};

HB_HD void hb_flag(int *err, int bit)
{
#ifdef __CUDA_ARCH__
	atomicOr(err, bit);
#else
	*err |= bit;
#endif
}
HB_HD int hb_fc_shift_(uint64_t e) { const int32_t v = (int32_t)((uint32_t)e >> 1); return (e & 1) ? -v : v; }
HB_HD int64_t hb_y_start_offset(int64_t x_start, const uint64_t *fc, uint32_t n, int *bad)
{ // y_start_offset, Hash_Table.h:165-189
	if (x_start == (int64_t)(fc[n - 1] >> 32)) return hb_fc_shift_(fc[n - 1]);
	uint32_t i = 0;
	for (; i < n; i++) if (x_start < (int64_t)(fc[i] >> 32)) break;
	if (i == 0 || i == n) { *bad = 1; return 0; }
	return hb_fc_shift_(fc[i - 1]);
}
HB_HD int hb_init_waln(int64_t err, int64_t s, int64_t l, int64_t w_l, int64_t *aux_beg, int64_t *aux_end, int64_t *r_s, int64_t *r_l)
{ // init_waln, Correct.cpp:764-780
	*aux_beg = *aux_end = *r_s = *r_l = -1;
	if (s < 0 || s >= l || (l - s + 2 * err + HB_THRE_MAX) < w_l) return 0;
	*aux_beg = *aux_end = 0;
	*r_s = s - err;
	*r_l = l - *r_s; if (*r_l > w_l) *r_l = w_l;
	*aux_end = w_l - *r_l;
	if (*r_s < 0) { *aux_beg = -*r_s; *r_s = 0; *r_l -= *aux_beg; }
	return 1;
}

HB_HD void hb_ez_push_trace(EcEz &ez, uint32_t c, uint32_t len)
{ // push_trace, Levenshtein_distance.h:522-531
	c <<= 14;
	while (len >= 0x3fff) { if (ez.cn < HB_EC_CIG_TMP) ez.cig[ez.cn] = (uint16_t)(c + 0x3fff); ez.cn++; len -= 0x3fff; }
	if (len) { if (ez.cn < HB_EC_CIG_TMP) ez.cig[ez.cn] = (uint16_t)(c + len); ez.cn++; }
}

HB_HD void hb_ez_gen_trace(EcEz &ez, int32_t ptrim)
{ // gen_trace(ez, ptrim, reverse = 1), Levenshtein_distance.h:903-985 with one word per column
	if (ez.err > ez.thre) return;
	ez.cn = 0;
	int32_t V, H, D, mn, cur = ez.err, tn = ez.te + 1 - ez.ts, pn = tn + (ez.thre << 1), bd = (ez.thre << 1) + 1;
	int32_t poff = ez.pe, sft = bd - (pn - ez.pe - ptrim), i = tn, low = bd - 1, d = 0, pd = -1, pdn = 0;
	while (i > 0 && cur > 0) {
		const uint64_t *c5 = ez.path + (size_t)(i - 1) * 5; // D0, VP, VN, HP, HN of column i-1
		D = cur - (int32_t)((~(c5[0] >> sft)) & 1ULL); d = 0; mn = D;
		if (sft != low) { H = cur + (int32_t)((c5[4] >> sft) & 1ULL) - (int32_t)((c5[3] >> sft) & 1ULL); if (H + 1 == cur && H <= mn) { mn = H; d = 3; } }
		if (sft != 0) { V = cur + (int32_t)((c5[2] >> (sft - 1)) & 1ULL) - (int32_t)((c5[1] >> (sft - 1)) & 1ULL); if (V + 1 == cur && V <= mn) { mn = V; d = 2; } }
		if (d == 0) { if (D != cur) d = 1; i--; poff--; }
		else if (d == 2) { sft--; poff--; }
		else { i--; sft++; }
		if (d == pd) pdn++;
		else { if (pdn > 0) hb_ez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn); pd = d; pdn = 1; }
		cur = mn;
	}
	if (i > 0) {
		d = 0; poff -= i;
		if (d == pd) pdn += i;
		else { if (pdn > 0) hb_ez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	poff++;
	if (ez.ps < 0 || ez.ps >= ez.pl) ez.ps = poff;
	else if (poff > ez.ps) {
		d = 2; i = poff - ez.ps;
		if (d == pd) pdn += i;
		else { if (pdn > 0) hb_ez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	if (pdn > 0) hb_ez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn);
	const int32_t n = ez.cn < HB_EC_CIG_TMP ? ez.cn : HB_EC_CIG_TMP;
	for (int32_t k = 0; k < (n >> 1); k++) { const uint16_t t = ez.cig[k]; ez.cig[k] = ez.cig[n - k - 1]; ez.cig[n - k - 1] = t; }
}

// ed_band_cal_semi_64_w_absent_diag[_trace] (Levenshtein_distance.h:3727-3776 / 3778-3856): pattern = target[ps0, ps0+pn)
// on the overlap's strand, text = query[qs0, qs0+tn).  TRACE = false: plain version (sets err / pe).  TRACE = true:
// when ez.err <= thre on entry the end point is known and only the path + cigar are produced.
template <bool TRACE>
HB_HD void hb_ez_semi64(const RdView &T, int64_t ps0, int32_t pn, const RdView &Q, int64_t qs0, int32_t tn, int32_t thre, int32_t abs_diag, EcEz &ez)
{
	if (TRACE) {
		ez.cn = 0;
		if (ez.err > thre) { ez.thre = thre; ez.err = INT32_MAX; ez.pl = pn; ez.tl = tn; ez.ps = ez.pe = -1; ez.ts = 0; ez.te = tn - 1; }
		else if (ez.err == 0) { hb_ez_push_trace(ez, 0, (uint32_t)(ez.te + 1 - ez.ts)); ez.ps = ez.pe - (ez.te - ez.ts); return; }
	} else { ez.thre = thre; ez.err = INT32_MAX; ez.pl = pn; ez.tl = tn; ez.ps = ez.pe = -1; ez.ts = 0; ez.te = tn - 1; }
	uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN, X, D0, HN, HP, mm;
	int32_t bd, i, err = abs_diag, i_bd, tn0 = tn - 1, cut = thre + (thre << 1), c;
	if (pn > tn + cut || tn > pn + cut) return;
	bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn;
	for (i = 0, mm = 1ULL << abs_diag; i < bd; i++) { Peq[T.at(ps0 + i)] |= mm; mm <<= 1; }
	i_bd = (thre << 1) - abs_diag; VN = (1ULL << abs_diag) - 1;
	Peq[4] = 0; mm = 1ULL << (thre << 1);
	for (i = 0; i <= tn0; i++) {
		X = Peq[Q.at(qs0 + i)] | VN;
		D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0; HP = VN | ~(VP | D0);
		X = D0 >> 1;
		VN = X & HP; VP = HN | ~(X | HP);
		if (!(D0 & 1ULL)) { ++err; if (err > cut) return; }
		if (i < tn0) {
			Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
			++i_bd; c = 4;
			if (i_bd < pn) c = T.at(ps0 + i_bd);
			if (c < 4) Peq[c] |= mm;
		}
		if (TRACE) { uint64_t *c5 = ez.path + (size_t)i * 5; c5[0] = D0; c5[1] = VP; c5[2] = VN; c5[3] = HP; c5[4] = HN; }
	}
	if (!TRACE || ez.err > thre) {
		int32_t site = tn - 1 - abs_diag, ai = pn - tn + abs_diag, uge = INT32_MAX;
		for (i = 0; site < 0 && i < ai; i++, site++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
		if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site; }
		site -= i;
		while (i < ai) {
			err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); ++i;
			if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site + i; }
			if (i == thre) uge = err;
		}
		if (uge <= thre && uge == ez.err) ez.pe = site + thre;
	}
	if (TRACE) hb_ez_gen_trace(ez, abs_diag);
}

// Reserve_Banded_BPM_Extension[_REV] (Levenshtein_distance.h:71-214 / 216-359).  pattern = 'N' x aux_beg + target[ts, ts+pri_l)
// + 'N' x aux_end (fill_subregion, Correct.cpp:270-277), text = query[qs, qs+t_length).  The reference's Peq table is indexed
// by character and only its A/C/G/T rows are shifted: row 4 is the 'N' row, which keeps every bit it receives.
HB_HD void hb_get_error(int t_length, int errthold, int init_err, uint64_t VP, uint64_t VN, unsigned int *return_err, int *back_site)
{ // get_error, Levenshtein_distance.h:20-69
	*return_err = (unsigned int)-1;
	int site = t_length - 1, return_site = -1, available_i = 2 * errthold, i = 0; unsigned int ungap_error = (unsigned int)-1;
	if (init_err <= errthold && (unsigned int)init_err <= *return_err) { *return_err = (unsigned int)init_err; return_site = site; }
	while (i < available_i) {
		init_err += (int)((VP >> i) & 1ULL); init_err -= (int)((VN >> i) & 1ULL); ++i;
		if (init_err <= errthold && (unsigned int)init_err <= *return_err) { *return_err = (unsigned int)init_err; return_site = site + i; }
		if (i == errthold) ungap_error = (unsigned int)init_err;
	}
	if (ungap_error <= (unsigned int)errthold && ungap_error == *return_err) return_site = site + errthold;
	*back_site = return_site;
}
HB_HD void hb_bpm_extension(const RdView &T, int64_t ts, int64_t pri_l, int64_t aux_beg, int p_length, const RdView &Q, int64_t qs, int t_length, int errthold, int rev,
                            unsigned int *return_err, int *return_p_end, int *return_t_end)
{
	*return_err = (unsigned int)-1; *return_p_end = -1; *return_t_end = -1;
	uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN = 0, X, D0, HN, HP, Mask = 1ULL << (errthold << 1), b = 1;
	unsigned int line_error = (unsigned int)-1; int return_site, band_length = (errthold << 1) + 1, i, err = 0, i_bd = errthold << 1, last_high = errthold << 1;
	auto pat = [&](int j) -> int { const int64_t k = (rev ? p_length - j - 1 : j) - aux_beg; return (k < 0 || k >= pri_l) ? 4 : T.at(ts + k); };
	auto txt = [&](int j) -> int { return Q.at(qs + (rev ? t_length - j - 1 : j)); };
	for (i = 0; i < band_length; i++) { Peq[pat(i)] |= b; b <<= 1; }
	Peq[4] = 0;
	for (i = 0; i < t_length; i++) {
		X = Peq[txt(i)] | VN;
		D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0; HP = VN | ~(VP | D0);
		X = D0 >> 1;
		VN = X & HP; VP = HN | ~(X | HP);
		if (!(D0 & 1ULL)) { ++err; if (err - last_high > errthold) return; }
		hb_get_error(i + 1, errthold, err, VP, VN, &line_error, &return_site);
		if (line_error != (unsigned int)-1) {
			*return_t_end = rev ? t_length - i - 1 : i;
			*return_p_end = rev ? p_length - return_site - 1 : return_site;
			*return_err = line_error;
		}
		if (i == t_length - 1) break;
		Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
		++i_bd;
		Peq[pat(i_bd)] |= Mask;
	}
}

// ---- window grid and thresholds ---------------------------------------------------------------------------
HB_HD void hb_win_by_s(const EcZ &z, int64_t w_s, int64_t bs, int64_t *w_e)
{ // get_win_id_by_s, Correct.h:1306-1314
	const int64_t n_s = (z.x_pos_s / bs) * bs, wid = (w_s - n_s) / bs;
	*w_e = n_s + (wid + 1) * bs - 1; if (*w_e > z.x_pos_e) *w_e = z.x_pos_e;
}
HB_HD void hb_win_by_e(const EcZ &z, int64_t w_e, int64_t bs, int64_t *w_s)
{ // get_win_id_by_e, Correct.h:1317-1325
	const int64_t n_s = (z.x_pos_s / bs) * bs, wid = (w_e - n_s) / bs;
	*w_s = n_s + wid * bs; if (*w_s < z.x_pos_s) *w_s = z.x_pos_s;
}
HB_HD int64_t hb_adj_thre(int64_t t, int64_t len) { return (t == 0 && len >= 4) ? 1 : t; } // Adjust_Threshold, Correct.h:46
HB_HD int64_t hb_ext_thres(int64_t len, double e_rate, int64_t bs, int64_t block_err)
{ // double_error_threshold(get_init_err_thres(len, e_rate, bs, block_err), len), Correct.cpp:1042-1049, 917-934
	int64_t t;
	if (len >= bs) t = block_err;
	else { t = (int64_t)((double)len * e_rate); t = hb_adj_thre(t, len); if (t > HB_THRE_MAX) t = HB_THRE_MAX; }
	int pre = (int)hb_adj_thre((int)t, (int)len), th = pre * 2;
	if (len >= 300 && th < HB_THRE_MAX) th = HB_THRE_MAX;
	if (th > HB_THRE_MAX) th = HB_THRE_MAX;
	return th;
}

// push_wcigar, Correct.cpp:4050-4055: the window's cigar goes to the shared pool (bump allocation; a cigar that is
// replaced by recal_boundary_exz simply leaves its first copy behind)
HB_HD void hb_push_wcigar(EcCtx &C, hb_wl_t *p)
{
	const int32_t n = C.ez.cn;
	if (n > HB_EC_CIG_TMP) { hb_flag(C.err, 64); p->cidx = 0; p->clen = 0; return; }
#ifdef __CUDA_ARCH__
	const unsigned long long o = atomicAdd(C.pool_used, (unsigned long long)n);
#else
	const unsigned long long o = *C.pool_used; *C.pool_used += (unsigned long long)n;
#endif
	p->clen = (uint32_t)n; p->cidx = (uint32_t)o;
	if (o + (unsigned long long)n > C.pool_cap) { p->cidx = 0; return; } // the host sees pool_used > pool_cap and reruns with a larger pool
	for (int32_t k = 0; k < n; k++) C.pool[o + k] = C.ez.cig[k];
}

HB_HD int hb_recal_boundary(EcCtx &C, const EcZ &z, int64_t qs, int64_t ql0, int64_t tl0, int64_t thres, int64_t toff, int64_t ts0, int64_t te0, int64_t err0,
                            int64_t *ts_r, int64_t *aux_beg_r, int64_t *aux_end_r)
{ // recal_boundary_exz, Correct.cpp:2429-2469
	int64_t ts, aux_beg, aux_end, t_pri_l, aln_l = ql0 + (thres << 1);
	if (ts0 == 0) ts = toff;
	else if (te0 + 1 == tl0) ts = toff + te0 - ql0 + 1;
	else return 0;
	if (!hb_init_waln(thres, ts, C.t.len, aln_l, &aux_beg, &aux_end, &ts, &t_pri_l)) return 0;
	if (ts == toff && tl0 == t_pri_l) return 0;
	C.ez.err = INT32_MAX;
	hb_ez_semi64<true>(C.t, ts, (int32_t)t_pri_l, C.q, qs, (int32_t)ql0, (int32_t)thres, (int32_t)aux_beg, C.ez);
	if (C.ez.err <= C.ez.thre && C.ez.err < err0) { *aux_beg_r = aux_beg; *aux_end_r = aux_end; *ts_r = ts; return 1; }
	return 0;
}

HB_HD uint32_t hb_aln_wlst_adv(EcCtx &C, EcZ &z, int64_t max_err, int64_t qs, int64_t qe, int64_t t_s, int is_cigar)
{ // aln_wlst_adv_exz, Correct.cpp:4057-4127
	const int64_t ql = qe + 1 - qs; int64_t aux_beg, aux_end, t_pri_l;
	const int64_t thres = hb_ext_thres(ql, C.e_rate, C.w_l, max_err), aln_l = ql + (thres << 1);
	EcEz &ez = C.ez;
	if (!hb_init_waln(thres, t_s, C.t.len, aln_l, &aux_beg, &aux_end, &t_s, &t_pri_l)) return 0;
	if (t_pri_l + thres < ql) return 0;
	const int64_t tl = t_pri_l;
	if (is_cigar) { ez.err = INT32_MAX; hb_ez_semi64<true>(C.t, t_s, (int32_t)tl, C.q, qs, (int32_t)ql, (int32_t)thres, (int32_t)aux_beg, ez); }
	else { hb_ez_semi64<false>(C.t, t_s, (int32_t)tl, C.q, qs, (int32_t)ql, (int32_t)thres, (int32_t)aux_beg, ez); ez.ps = 0; }
	if (ez.err > ez.thre) return 0;
	hb_wl_t *p = &z.w[z.wn++];
	p->x_start = (int32_t)qs; p->x_end = (int32_t)qe;
	p->y_start = (int32_t)(t_s + ez.ps); p->y_end = (int32_t)(t_s + ez.pe);
	p->error = (int16_t)ez.err; p->cidx = p->clen = 0;
	if (is_cigar) {
		hb_push_wcigar(C, p);
		if ((ez.pe + 1 == tl || ez.ps == 0) && ez.err > 0) {
			if (hb_recal_boundary(C, z, qs, ql, tl, thres, t_s, ez.ps, ez.pe, ez.err, &t_s, &aux_beg, &aux_end)) {
				hb_push_wcigar(C, p);
				p->y_start = (int32_t)(t_s + ez.ps); p->y_end = (int32_t)(t_s + ez.pe); p->error = (int16_t)ez.err;
			}
		}
	}
	p->extra_begin = (int16_t)aux_beg; p->extra_end = (int16_t)aux_end; p->error_threshold = (int16_t)thres;
	z.align_length += ql;
	return 1;
}

HB_HD uint32_t hb_gen_backtrace_adv(EcCtx &C, EcZ &z, hb_wl_t *p)
{ // gen_backtrace_adv_exz, Correct.cpp:12563-12641
	if (p->error < 0 || p->y_end < 0) return 0;
	const int64_t qs = p->x_start, ql = (int64_t)p->x_end + 1 - qs, thres = p->error_threshold, aln_l = ql + (thres << 1);
	int64_t ts = p->y_start, aux_beg = p->extra_begin, aux_end = p->extra_end, t_pri_l;
	EcEz &ez = C.ez;
	if (aux_end >= 0) t_pri_l = aln_l - aux_beg - aux_end;
	else { t_pri_l = ts + aln_l - aux_beg; if (t_pri_l > (int64_t)C.t.len) t_pri_l = C.t.len; t_pri_l -= ts; }
	const int64_t tl = t_pri_l;
	ez.ts = 0; ez.te = p->x_end - p->x_start; ez.tl = (int32_t)ql;
	ez.ps = -1; ez.pe = p->y_end - p->y_start; ez.pl = (int32_t)tl;
	ez.err = p->error; ez.thre = p->error_threshold;
	hb_ez_semi64<true>(C.t, ts, (int32_t)tl, C.q, qs, (int32_t)ql, (int32_t)thres, (int32_t)aux_beg, ez);
	if (ez.err <= ez.thre) {
		p->y_start = (int32_t)(ts + ez.ps); p->y_end = (int32_t)(ts + ez.pe); p->error = (int16_t)ez.err;
		hb_push_wcigar(C, p);
		if ((ez.pe + 1 == tl || ez.ps == 0) && ez.err > 0) {
			if (hb_recal_boundary(C, z, qs, ql, tl, thres, ts, ez.ps, ez.pe, ez.err, &ts, &aux_beg, &aux_end)) {
				hb_push_wcigar(C, p);
				p->y_start = (int32_t)(ts + ez.ps); p->y_end = (int32_t)(ts + ez.pe); p->error = (int16_t)ez.err;
			}
		}
		p->extra_begin = (int16_t)aux_beg; p->extra_end = (int16_t)aux_end;
		return 1;
	}
	p->error = -1;
	return 0;
}

// push_hc_wlst_exz, Correct.cpp:12776-12835 (force_aln = 0), for the aligned window `wr` of the window pass
HB_HD uint32_t hb_push_hc_wlst(EcCtx &C, EcZ &z, const hb_win_t &wr, int64_t tl)
{
	hb_wl_t p; int64_t w_e, w_s, ce = (int64_t)wr.q_s - 1, cs = z.x_pos_s, toff, ys;
	p.x_start = wr.q_s; p.x_end = wr.q_e; p.y_start = wr.t_s; p.y_end = wr.t_s + wr.pe; p.error = (int16_t)wr.err;
	p.extra_begin = (int16_t)wr.aux_beg; p.extra_end = (int16_t)wr.aux_end; p.error_threshold = (int16_t)wr.thre; p.cidx = p.clen = 0;
	if (z.wn > 0) { // forward from the end of the previous window
		w_e = z.w[z.wn - 1].x_end; toff = (int64_t)z.w[z.wn - 1].y_end + 1;
		while (w_e < ce && toff < tl) {
			w_s = w_e + 1; hb_win_by_s(z, w_s, C.w_l, &w_e);
			if (hb_aln_wlst_adv(C, z, HB_THRE_MAX, w_s, w_e, toff, 0)) toff = (int64_t)z.w[z.wn - 1].y_end + 1;
			else break;
		}
		cs = (int64_t)z.w[z.wn - 1].x_end + 1;
	}
	const int32_t a_n = z.wn; w_s = wr.q_s;
	if (w_s > cs) { // backward from the traced start of this window
		hb_gen_backtrace_adv(C, z, &p);
		toff = (int64_t)p.y_start - 1;
		while (w_s > cs) {
			w_e = w_s - 1; hb_win_by_e(z, w_e, C.w_l, &w_s); ys = toff + 1 - (w_e + 1 - w_s);
			if (ys >= 0 && hb_aln_wlst_adv(C, z, HB_THRE_MAX, w_s, w_e, ys, 1)) toff = (int64_t)z.w[z.wn - 1].y_start - 1;
			else break;
		}
	}
	z.align_length += (int64_t)wr.q_e + 1 - wr.q_s;
	const int64_t ovl = z.x_pos_e + 1 - z.x_pos_s, ualn = ((int64_t)wr.q_e + 1 - z.x_pos_s) - z.align_length, aln = ovl - ualn;
	if (!(aln > 0 && (double)ovl * HB_OVLP_CUT <= (double)aln)) { z.w[z.wn++] = p; return 0; }
	for (int32_t i = a_n, j = z.wn - 1; i < j; i++, j--) { const hb_wl_t t = z.w[i]; z.w[i] = z.w[j]; z.w[j] = t; }
	z.w[z.wn++] = p;
	return 1;
}

HB_HD uint32_t hb_ed_cut(EcCtx &C, const EcZ &z, int64_t qs, int64_t qe, int64_t t_s, int64_t max_err, uint32_t aln_dir, int64_t *r_err, int64_t *aln_qlen)
{ // ed_cut, Correct.cpp:13063-13104
	*aln_qlen = 0; *r_err = INT32_MAX;
	const int64_t ql = qe + 1 - qs, thres = hb_ext_thres(ql, C.e_rate, C.w_l, max_err), aln_l = ql + (thres << 1);
	int64_t aux_beg, aux_end, t_pri_l; unsigned int error; int t_end, q_end;
	if (!hb_init_waln(thres, t_s, C.t.len, aln_l, &aux_beg, &aux_end, &t_s, &t_pri_l)) return 0;
	hb_bpm_extension(C.t, t_s, t_pri_l, aux_beg, (int)aln_l, C.q, qs, (int)ql, (int)thres, aln_dir ? 1 : 0, &error, &t_end, &q_end);
	if (t_end != -1 && q_end != -1) *aln_qlen = aln_dir ? ql - q_end : q_end + 1;
	*r_err = error;
	return *aln_qlen == 0 ? 0 : 1;
}

HB_HD int64_t hb_gen_extend_err_0(EcCtx &C, EcZ &z, int64_t max_err, int64_t qs, int64_t qe, int64_t pk, int *bad)
{ // gen_extend_err_0_exz, Correct.cpp:13224-13283
	int64_t tot_e = 0, ts, di[2], al[2], tb[2]; const int64_t an = z.wn, ql = qe + 1 - qs;
	ts = (qs - z.x_pos_s) + z.y_pos_s; ts += hb_y_start_offset(qs, z.fc, z.fc_n, bad);
	di[0] = di[1] = al[0] = al[1] = 0; tb[0] = tb[1] = -1;
	if (pk > 0 && qs == (int64_t)z.w[pk].x_end + 1) {
		if (z.w[pk].clen == 0) hb_gen_backtrace_adv(C, z, &z.w[pk]);
		tb[0] = (int64_t)z.w[pk].y_end + 1;
	}
	if (pk + 1 < an && qe + 1 == (int64_t)z.w[pk + 1].x_start) {
		if (z.w[pk + 1].clen == 0) hb_gen_backtrace_adv(C, z, &z.w[pk + 1]);
		tb[1] = (int64_t)z.w[pk + 1].y_start - ql;
	}
	if (tb[0] == -1 && tb[1] == -1) tb[0] = tb[1] = ts;
	else if (tb[0] == -1 && tb[1] != -1) tb[0] = tb[1];
	else if (tb[1] == -1 && tb[0] != -1) tb[1] = tb[0];
	if (tb[0] != -1) { if (!hb_ed_cut(C, z, qs, qe, tb[0], max_err, 0, &di[0], &al[0])) { di[0] = ql; al[0] = 0; } }
	if (tb[1] != -1) { if (!hb_ed_cut(C, z, qs, qe, tb[1], max_err, 1, &di[1], &al[1])) { di[1] = ql; al[1] = 0; } }
	if (al[0] && al[1]) {
		if (al[0] + al[1] <= ql) tot_e += di[0] + di[1] + ql - (al[0] + al[1]);
		else { const double rr = (double)ql / (double)(al[0] + al[1]); tot_e = (int64_t)((double)tot_e + (double)(di[0] + di[1]) * rr); }
	} else if (!al[0] && !al[1]) tot_e += ql;
	else if (al[0]) tot_e += di[0] + (ql - al[0]);
	else tot_e += di[1] + (ql - al[1]);
	return tot_e;
}

HB_HD double hb_gen_extend_err(EcCtx &C, EcZ &z, double e_max, int64_t max_err, int64_t *r_e, int *bad)
{ // gen_extend_err_exz, Correct.cpp:13400-13440 (sec_check = 0)
	const int64_t ovl = z.x_pos_e + 1 - z.x_pos_s, an = z.wn; int64_t k, ce, tot_l = 0, tot_e = 0, ws, we;
	*r_e = INT64_MAX;
	for (k = an - 1, ce = z.x_pos_e; k >= 0; k--) {
		tot_l += (int64_t)z.w[k].x_end + 1 - z.w[k].x_start;
		tot_e += z.w[k].error;
		we = z.w[k].x_end;
		while (we < ce) {
			ws = we + 1; hb_win_by_s(z, ws, C.w_l, &we);
			tot_l += we + 1 - ws;
			tot_e += hb_gen_extend_err_0(C, z, max_err, ws, we, k, bad);
			if (e_max > 0 && (double)tot_e > (double)ovl * e_max) return 1.7976931348623157e308;
		}
		ce = (int64_t)z.w[k].x_start - 1;
		if (e_max > 0 && (double)tot_e > (double)ovl * e_max) return 1.7976931348623157e308;
	}
	if (ce >= z.x_pos_s) {
		we = z.x_pos_s - 1;
		while (we < ce) {
			ws = we + 1; hb_win_by_s(z, ws, C.w_l, &we);
			tot_l += we + 1 - ws;
			tot_e += hb_gen_extend_err_0(C, z, max_err, ws, we, k, bad);
			if (e_max > 0 && (double)tot_e > (double)ovl * e_max) return 1.7976931348623157e308;
		}
	}
	*r_e = tot_e;
	return (double)tot_e / (double)tot_l;
}

// One overlap: the window records of k_windows (win[0..nw), in window order) -> window list + acceptance.
// out->st: 0 = rejected by the window pass, 1 = aligned but rr > e_rate, 2 = accepted (re = error estimate).
HB_HD void hb_ec_overlap_A(EcCtx &C, const hb_chain_t &c, const uint64_t *fc, const hb_win_t *win, int32_t nw, hb_wl_t *wl, hb_aln_t *out)
{
	EcZ z; z.x_pos_s = c.x_pos_s; z.x_pos_e = c.x_pos_e; z.y_pos_s = c.y_pos_s; z.y_id = c.y_id; z.rev = c.y_pos_strand; z.fc = fc; z.fc_n = c.fc_n;
	z.align_length = 0; z.w = wl; z.wn = 0;
	const int64_t tl = C.t.len; uint32_t ok = 1; int bad = 0;
	for (int32_t k = 0; k < nw; k++) { // align_hc_ed_post_extz, Correct.cpp:12951-13011
		const hb_win_t wr = win[k];
		if (wr.t_pri_l < 0 || wr.err > wr.thre) continue;
		if (!hb_push_hc_wlst(C, z, wr, tl)) { ok = 0; break; }
	}
	if (ok) { const int64_t ovl = z.x_pos_e + 1 - z.x_pos_s; if (!(z.align_length > 0 && (double)ovl * HB_OVLP_CUT <= (double)z.align_length)) ok = 0; }
	int64_t re = INT64_MAX; double rr = 1.7976931348623157e308;
	if (ok) rr = hb_gen_extend_err(C, z, C.e_rate * 1.5 + 0.000001, HB_THRE_MAX, &re, &bad);
	if (bad) hb_flag(C.err, 32);
	out->st = !ok ? 0 : (rr > C.e_rate ? 1 : 2); out->align_length = (uint32_t)z.align_length; out->rr = rr; out->re = re; out->w_n = (uint32_t)z.wn;
}
