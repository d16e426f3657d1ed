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
#include "hb_warp.cuh"

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
// the same through a cursor that keeps the 64-bit word (32 bases) it last touched: the aligner walks both reads base by base, in one direction
// (a read's packed slot is 32-byte aligned and padded, so the word holding its last base can be loaded whole)
struct RdCur { uint64_t w; uint32_t idx; };
HB_HD int hb_rd_at(const RdView &v, RdCur &c, int64_t j)
{
	const uint32_t fp = v.rev ? (uint32_t)(v.len - 1 - j) : (uint32_t)j;
	if (v.nn) { uint32_t lo = 0, hi = v.nn; while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (v.npos[mid] < fp) lo = mid + 1; else hi = mid; } if (lo < v.nn && v.npos[lo] == fp) return 4; }
	const uint32_t wi = fp >> 5;
	if (wi != c.idx) { c.w = ((const uint64_t *)v.p)[wi]; c.idx = wi; }
	const int b = (int)((c.w >> ((((fp & 31) >> 2) << 3) + ((3 - (fp & 3)) << 1))) & 3);
	return v.rev ? 3 - b : b;
}
HB_HD RdView hb_rd_view(const DevReads &R, uint64_t id, uint32_t rev)
{
	RdView v; v.p = R.packed + R.off[id]; v.npos = R.npos + R.noff[id]; v.nn = (uint32_t)(R.noff[id + 1] - R.noff[id]); v.len = R.len[id]; v.rev = rev;
	return v;
}

struct OvDesc { uint32_t read, slot, nw, pad; uint64_t w0; }; // an overlap of the batch: batch-local read, chain slot, number of windows, first window

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

// =============================================================================================================
// step B: base-level CIGAR of an accepted overlap — gen_hc_fast_cigar (Correct.cpp:25137 -> 17813, row a10):
// return_t_chain (22997: lchain_refine of the chain's anchors), hc_ovlp_base_direct (17425: one alignment per
// inter-anchor segment, exact shortcuts first, then hc_aln_exz_adv_hc 16178 with its threshold escalation
// estimate -> len*e_rate -> x2 -> 0.51*len -> max), push_alnw (15988: consecutive aligned segments are fused into one
// window whose cigar grows), update_overlap_region (17249).  Alignment = multi-word banded Myers with traceback
// (ed_band_cal_{global,extension_0,extension_1,semi}_infi_w_trace, Levenshtein_distance.h:2516/2694/2823/3020).
// One thread per overlap; scratch sizes are launch parameters, an overlap that does not fit is reported as deferred
// (st = -1) and re-run by a second launch with few threads and large scratch.
// On the GPU the three pieces of the step run as three kernels: prep (thread / overlap), segment (thread / inter-anchor segment — the
// alignments are independent of each other), merge (thread / overlap).
// =============================================================================================================
#define HB_MAX_SIN_L 10000 // Levenshtein_distance.h:757
#define HB_MAX_SIN_E 2047  // Levenshtein_distance.h:756
#define HB_FORCE_SIN_L 512 // Levenshtein_distance.h:758
#define HB_MW_MAXW 64      // words of a band of 2*2047+1 bits
#define HB_RC_SPARE_WIN 8   // windows an overlap's list may grow by in the re-seeding rescue (hb_ecrechain.cuh); more: the overlap is reported

struct MwEz {
	int32_t ps, pe, pl, ts, te, tl, thre, err, nword;
	uint16_t *cig; int32_t cn, ccap;      // cigar of the last alignment
	uint64_t *path; uint64_t pcap, pn;    // 5*nword words per column; one-word bands (compact = 1): 2 header words (initial VP, VN) + 3 per column (D0, VP, VN)
	int32_t compact;                      // trace layout: 0 = 5 words per band word and column, 1 = one-word band (3 per column), 2 = hb_mw_align_w's (3 per band word and column), 3 = one-word band of <= 21 bits (1 per column)
	int32_t warp;                         // != 0: the calling WARP runs the aligner together (hb_mwalign_w.cuh); every lane holds the same MwEz
	uint64_t *vec; int32_t vstride;       // 11 vectors of vstride words: Peq[0..4], VP, VN, X, D0, HN, HP
	int ovf;                              // scratch too small: the unit is deferred to a launch with more scratch
};
// result of one segment's alignment as push_alnw consumes it
struct AlnRes { int32_t ts, te, ps, pe, err, cn; const uint16_t *cig; };
struct EcBCtx {
	RdView q, t; int64_t ql, tl; double e_rate; int64_t w_l;
	MwEz ez;
	hb_wl_t *aw; int32_t awn, awcap;      // aux_o->w_list of the overlap
	uint16_t *wc; int32_t wcn, wccap;     // cigar of the window under construction (aux_o->w_list.c tail), flushed to the pool when the window closes
	int32_t open;                          // index of the window whose cigar is in wc (-1: none)
	int32_t do_gaps; int64_t re_A, gap_re; // step C (reassign_gaps) applied when a window closes; errors removed by it
	int32_t no_myers;                      // segment pre-pass: stop (status 5) where an alignment would start
	uint16_t *pool; unsigned long long *pool_used; uint64_t pool_cap;
	uint16_t *gout; int32_t gcap;          // output of reassign_gaps at a flush; 0: ez.cig (only where no alignment result is pending in it when a window closes)
	int bad;
};

HB_HD void hb_mez_push_trace(MwEz &ez, uint32_t c, uint32_t len)
{
	c <<= 14;
	while (len >= 0x3fff) { if (ez.cn < ez.ccap) ez.cig[ez.cn] = (uint16_t)(c + 0x3fff); else ez.ovf = 1; ez.cn++; len -= 0x3fff; }
	if (len) { if (ez.cn < ez.ccap) ez.cig[ez.cn] = (uint16_t)(c + len); else ez.ovf = 1; ez.cn++; }
}
HB_HD int hb_mw_bit(const uint64_t *x, int32_t b) { return (int)((x[b >> 6] >> (b & 63)) & 1ULL); }
HB_HD void hb_mw_set_lsub(uint64_t *x, int32_t l, int32_t nw)
{ // w_infi_set_bit_lsub, Levenshtein_distance.h:2136
	for (int32_t k = 0; k < nw; k++) x[k] = k < (l >> 6) ? ~0ULL : 0ULL;
	if (l & 63) x[l >> 6] = (1ULL << (l & 63)) - 1;
}

HB_HD_NI void hb_mw_gen_trace(MwEz &ez, int32_t ptrim, int reverse)
{ // gen_trace, Levenshtein_distance.h:903-985
	if (ez.err > ez.thre) return;
	ez.cn = 0;
	int32_t V, H, D, mn, cur = ez.err, tn = ez.te + 1 - ez.ts, pn = tn + (ez.thre << 1), bd = (ez.thre << 1) + 1;
	const int32_t bs = ez.compact ? 0 : (int32_t)(ez.pn / (uint64_t)tn), bbs = bs / 5;
	int32_t poff = ez.pe, sft = bd - (pn - ez.pe - ptrim), i = tn, low = bd - 1, d = 0, pd = -1, pdn = 0, c3_i = -1; uint64_t c3_w = 0, c3_pw = 0, c3_nx = 0;
	while (i > 0 && cur > 0) {
		const uint64_t *D0 = ez.path + (size_t)(i - 1) * bs, *VP = D0 + bbs, *VN = VP + bbs, *HP = VN + bbs, *HN = HP + bbs;
		if (ez.compact == 2) { // rows of 3 * nword words (D0 | VP | VN) behind a header of the initial VP | VN: HN / HP from the row before, bit by bit
			const int32_t nw = ez.nword; const uint64_t *row = ez.path + 2 * (size_t)nw + (size_t)(i - 1) * 3 * nw, *pvp = i > 1 ? row - 2 * nw : ez.path, *pvn = pvp + nw;
			const int d0b = hb_mw_bit(row, sft);
			D = cur - (1 - d0b); d = 0; mn = D;
			if (sft != low) { const int vppb = hb_mw_bit(pvp, sft), hn = vppb & d0b, hp = hb_mw_bit(pvn, sft) | (1 - (vppb | d0b)); H = cur + hn - hp; if (H + 1 == cur && H <= mn) { mn = H; d = 3; } }
			if (sft != 0) { V = cur + hb_mw_bit(row + 2 * nw, sft - 1) - hb_mw_bit(row + nw, sft - 1); if (V + 1 == cur && V <= mn) { mn = V; d = 2; } }
		} else if (ez.compact == 3) { // one-word band of <= 21 bits: one trace word per column (D0 | VP << 21 | VN << 42), header = initial VP | VN << 21
			if (c3_i != i) { c3_w = ez.path[i]; c3_pw = ez.path[i - 1]; c3_nx = i > 1 ? ez.path[i - 2] : 0; c3_i = i; } // rows i, i-1 in registers, row i-2 already on its way
			const uint64_t w = c3_w, pw = c3_pw; // column i-1 sits at path[1 + (i-1)]
			const uint64_t d0 = w & 0x1fffffULL, vp = (w >> 21) & 0x1fffffULL, vn = (w >> 42) & 0x1fffffULL;
			const uint64_t vpp = i > 1 ? (pw >> 21) & 0x1fffffULL : pw & 0x1fffffULL, vnp = i > 1 ? (pw >> 42) & 0x1fffffULL : (pw >> 21) & 0x1fffffULL, hn = vpp & d0, hp = vnp | ~(vpp | d0);
			D = cur - (1 - (int32_t)((d0 >> sft) & 1ULL)); d = 0; mn = D;
			if (sft != low) { H = cur + (int32_t)((hn >> sft) & 1ULL) - (int32_t)((hp >> sft) & 1ULL); if (H + 1 == cur && H <= mn) { mn = H; d = 3; } }
			if (sft != 0) { V = cur + (int32_t)((vn >> (sft - 1)) & 1ULL) - (int32_t)((vp >> (sft - 1)) & 1ULL); if (V + 1 == cur && V <= mn) { mn = V; d = 2; } }
		} else if (ez.compact) { // one-word band: D0, VP, VN of the column; HN = VP' & D0, HP = VN' | ~(VP' | D0) with VP', VN' of the column before (header for column 0)
			const uint64_t *row = ez.path + 2 + (size_t)(i - 1) * 3, *prv = i > 1 ? row - 3 + 1 : ez.path;
			const uint64_t d0 = row[0], vp = row[1], vn = row[2], vpp = prv[0], vnp = prv[1], hn = vpp & d0, hp = vnp | ~(vpp | d0);
			D = cur - (1 - (int32_t)((d0 >> sft) & 1ULL)); d = 0; mn = D;
			if (sft != low) { H = cur + (int32_t)((hn >> sft) & 1ULL) - (int32_t)((hp >> sft) & 1ULL); if (H + 1 == cur && H <= mn) { mn = H; d = 3; } }
			if (sft != 0) { V = cur + (int32_t)((vn >> (sft - 1)) & 1ULL) - (int32_t)((vp >> (sft - 1)) & 1ULL); if (V + 1 == cur && V <= mn) { mn = V; d = 2; } }
		} else {
		D = cur - (1 - hb_mw_bit(D0, sft)); d = 0; mn = D;
		if (sft != low) { H = cur + hb_mw_bit(HN, sft) - hb_mw_bit(HP, sft); if (H + 1 == cur && H <= mn) { mn = H; d = 3; } }
		if (sft != 0) { V = cur + hb_mw_bit(VN, sft - 1) - hb_mw_bit(VP, sft - 1); if (V + 1 == cur && V <= mn) { mn = V; d = 2; } }
		}
		if (d == 0) { if (D != cur) d = 1; i--; poff--; }
		else if (d == 2) { sft--; poff--; }
		else { i--; sft++; }
		if (ez.compact == 3 && c3_i == i + 1 && i > 0) { c3_w = c3_pw; c3_pw = c3_nx; c3_i = i; c3_nx = i > 1 ? ez.path[i - 2] : 0; } // the column moved back by one: shift the rows, fetch the next one early
		if (d == pd) pdn++;
		else { if (pdn > 0) hb_mez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn); pd = d; pdn = 1; }
		cur = mn;
	}
	if (i > 0) {
		d = 0; poff -= i;
		if (d == pd) pdn += i;
		else { if (pdn > 0) hb_mez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	poff++;
	if (ez.ps < 0 || ez.ps >= ez.pl) ez.ps = poff;
	else if (poff > ez.ps) {
		d = 2; i = poff - ez.ps;
		if (d == pd) pdn += i;
		else { if (pdn > 0) hb_mez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	if (pdn > 0) hb_mez_push_trace(ez, (uint32_t)pd, (uint32_t)pdn);
	if (reverse && !ez.ovf) for (int32_t k = 0, h = ez.cn >> 1; k < h; k++) { const uint16_t t = ez.cig[k]; ez.cig[k] = ez.cig[ez.cn - k - 1]; ez.cig[ez.cn - k - 1] = t; }
}

// mode: 0 global, 1 forward extension, 2 backward extension, 3 semi-global with abs_diag absent leading diagonals.
// pattern = target[ps0, ps0+pn) on the overlap's strand, text = query[qs0, qs0+tn).
HB_HD_NI void hb_mw_align(int mode, const RdView &T, int64_t ps0, int32_t pn, const RdView &Q, int64_t qs0, int32_t tn, int32_t thre, int32_t abs_diag, MwEz &ez)
{
	int32_t bd = (thre << 1) + 1; const int32_t nword = (bd >> 6) + ((bd & 63) ? 1 : 0), cut = thre + (thre << 1);
	int32_t i, err, i_bd, c, pidx = 0, tidx = 0, tmp_e = INT32_MAX, k, poff;
	ez.cn = 0; ez.thre = thre; ez.err = INT32_MAX; ez.pl = pn; ez.tl = tn;
	if (mode == 0) { ez.ps = ez.ts = 0; if (pn > tn + thre || tn > pn + thre) return; }
	else if (mode == 1) { ez.ps = ez.ts = 0; ez.pe = ez.te = -1; if (pn > tn + thre) pn = tn + thre; else if (tn > pn + thre) tn = pn + thre; }
	else if (mode == 2) { ez.ps = ez.ts = INT32_MAX; ez.pe = pn - 1; ez.te = tn - 1; if (pn > tn + thre) pn = tn + thre; else if (tn > pn + thre) tn = pn + thre; pidx = ez.pe; tidx = ez.te; }
	else { ez.ps = ez.pe = -1; ez.ts = 0; ez.te = tn - 1; if (pn > tn + cut || tn > pn + cut) return; }
	const int32_t tn0 = tn - 1, pe = pn - 1;
	ez.nword = nword;
	const bool pk21 = bd <= 21; // the whole band in 21 bits (thre <= 10: the usual segment): D0 | VP | VN of a column share ONE trace word
	if ((nword == 1 ? (pk21 ? 1 + (uint64_t)tn : 2 + 3 * (uint64_t)tn) : (uint64_t)nword * (uint64_t)tn * 5) > ez.pcap || nword > ez.vstride) { ez.ovf = 1; return; }
	ez.compact = nword == 1 ? (pk21 ? 3 : 1) : 0;
	if (nword == 1) { // the band fits one word (thre <= 31: the bulk of the segments): same algorithm with the vectors in registers
		uint64_t P0 = 0, P1 = 0, P2 = 0, P3 = 0, VP, VN, X, D0 = 0, HN = 0, HP = 0;
		RdCur curT, curQ; curT.idx = curQ.idx = 0xffffffffu; curT.w = curQ.w = 0;
		auto pch1 = [&](int32_t j) -> int { return hb_rd_at(T, curT, ps0 + (mode == 2 ? pidx - j : j)); };
		auto tch1 = [&](int32_t j) -> int { return hb_rd_at(Q, curQ, qs0 + (mode == 2 ? tidx - j : j)); };
		auto peq_or = [&](int cc, uint64_t m) { if (cc == 0) P0 |= m; else if (cc == 1) P1 |= m; else if (cc == 2) P2 |= m; else if (cc == 3) P3 |= m; };
		if (mode == 3) { VP = 0; VN = (1ULL << abs_diag) - 1; bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn; i_bd = abs_diag; for (i = 0; i < bd; i++, i_bd++) peq_or(pch1(i), 1ULL << i_bd); i_bd = (thre << 1) - abs_diag; err = abs_diag; }
		else { bd = thre + 1; bd = bd <= pn ? bd : pn; i_bd = thre; for (i = 0; i < bd; i++, i_bd++) peq_or(pch1(i), 1ULL << i_bd); i_bd = thre; err = thre; VN = (1ULL << thre) - 1; VP = ((1ULL << ((thre << 1) + 1)) - 1) ^ VN; }
		const uint64_t mm = 1ULL << (thre << 1);
		const uint64_t M21 = (1ULL << 21) - 1;
		uint64_t *pp = ez.path; // the trace cursor lives in a register for the loop (ez is a struct in memory: its fields would be reloaded every column)
		if (pk21) { *pp++ = (VP & M21) | (VN & M21) << 21; }
		else { pp[0] = VP; pp[1] = VN; pp += 2; } // HP / HN of a column follow from its D0 and the column before: not stored
		const int32_t ethre = ez.thre;
		for (i = 0; i <= tn0; i++) {
			const int tc = tch1(i);
			X = (tc == 0 ? P0 : tc == 1 ? P1 : tc == 2 ? P2 : tc == 3 ? P3 : 0ULL) | VN;
			D0 = ((VP + (X & VP)) ^ VP) | X;
			HN = VP & D0; HP = VN | ~(VP | D0);
			X = D0 >> 1;
			VN = X & HP; VP = HN | ~(X | HP);
			if (!(D0 & 1ULL)) { ++err; if (err > cut) return; }
			if (i < tn0) {
				if (mode == 1 || mode == 2) {
					poff = i - thre; k = i + thre - pe;
					if (k >= 0) {
						if (tmp_e == INT32_MAX) { tmp_e = err; for (k = 0; poff < pe; poff++, k++) { tmp_e += (int32_t)((VP >> k) & 1ULL); tmp_e -= (int32_t)((VN >> k) & 1ULL); } }
						else { k = (thre << 1) - k; if (k >= 0) { tmp_e += (int32_t)((HP >> k) & 1ULL); tmp_e -= (int32_t)((HN >> k) & 1ULL); } }
						if (tmp_e <= ethre && tmp_e < ez.err) { ez.err = tmp_e; if (mode == 1) { ez.pe = pe; ez.te = i; } else { ez.ps = pidx - pe; ez.ts = tidx - i; } }
					}
				}
				P0 >>= 1; P1 >>= 1; P2 >>= 1; P3 >>= 1;
				++i_bd;
				if (i_bd < pn) peq_or(pch1(i_bd), mm);
			}
			if (pk21) { *pp++ = (D0 & M21) | (VP & M21) << 21 | (VN & M21) << 42; }
			else { pp[0] = D0; pp[1] = VP; pp[2] = VN; pp += 3; }
		}
		ez.pn = (uint64_t)(pp - ez.path);
		if (mode == 0) {
			int32_t site = tn - 1 - thre; const int32_t ct = pn - 1;
			for (i = 0; site < ct; site++, i++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
			if (site == ct && err <= thre) { ez.err = err; ez.pe = pn - 1; ez.te = tn - 1; }
			hb_mw_gen_trace(ez, thre, 1);
		} else if (mode == 1 || mode == 2) {
			int32_t site = tn - 1 - thre; const int32_t ct = pn - 1;
			for (i = 0; site < ct; i++) {
				err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); site++;
				if (err <= thre && err < ez.err) { ez.err = err; if (mode == 1) { ez.pe = site; ez.te = tn - 1; } else { ez.ps = pidx - site; ez.ts = tidx + 1 - tn; } }
			}
			if (err <= thre && err < ez.err) { ez.err = err; if (mode == 1) { ez.pe = site; ez.te = tn - 1; } else { ez.ps = pidx - site; ez.ts = tidx + 1 - tn; } }
			if (mode == 1) hb_mw_gen_trace(ez, thre, 1);
			else { poff = ez.ps; ez.ps = pidx - ez.pe; ez.pe = pidx - poff; hb_mw_gen_trace(ez, thre, 0); poff = ez.ps; ez.ps = pidx - ez.pe; ez.pe = pidx - poff; }
		} else {
			int32_t site = tn - 1 - abs_diag, uge = INT32_MAX; const int32_t ai = pn - tn + abs_diag;
			for (i = 0; site < 0 && i < ai; i++, site++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
			if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site; }
			site -= i;
			while (i < ai) {
				err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); ++i;
				if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site + i; }
				if (i == thre) uge = err;
			}
			if (uge <= thre && uge == ez.err) ez.pe = site + thre;
			hb_mw_gen_trace(ez, abs_diag, 1);
		}
		return;
	}
	const int32_t VS = ez.vstride;
	uint64_t *Peq = ez.vec, *VP = ez.vec + 5 * VS, *VN = VP + VS, *X = VN + VS, *D0 = X + VS, *HN = D0 + VS, *HP = HN + VS;
	for (k = 0; k < 5 * VS; k++) Peq[k] = 0;
	auto pch = [&](int32_t j) -> int { return T.at(ps0 + (mode == 2 ? pidx - j : j)); };
	auto tch = [&](int32_t j) -> int { return Q.at(qs0 + (mode == 2 ? tidx - j : j)); };
	if (mode == 3) {
		for (k = 0; k < nword; k++) VP[k] = 0;
		hb_mw_set_lsub(VN, abs_diag, nword);
		bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn; i_bd = abs_diag;
		for (i = 0; i < bd; i++, i_bd++) { c = pch(i); Peq[c * VS + (i_bd >> 6)] |= 1ULL << (i_bd & 63); }
		i_bd = (thre << 1) - abs_diag; err = abs_diag;
	} else {
		bd = thre + 1; bd = bd <= pn ? bd : pn; i_bd = thre;
		for (i = 0; i < bd; i++, i_bd++) { c = pch(i); Peq[c * VS + (i_bd >> 6)] |= 1ULL << (i_bd & 63); }
		i_bd = thre; err = thre;
		hb_mw_set_lsub(VN, thre, nword); hb_mw_set_lsub(VP, (thre << 1) + 1, nword);
		for (k = 0; k < nword; k++) VP[k] ^= VN[k];
	}
	for (k = 0; k < nword; k++) Peq[4 * VS + k] = 0;
	ez.pn = 0;
	const int32_t Peq_i = (thre << 1) >> 6; const uint64_t Peq_m = 1ULL << ((thre << 1) & 63);
	for (i = 0; i <= tn0; i++) {
		{ // ed_infi_core, Levenshtein_distance.h:2148-2174
			const uint64_t *pq = Peq + tch(i) * VS; uint64_t ad = 0; int32_t w;
			for (w = 0; w < nword; w++) {
				const uint64_t x = pq[w] | VN[w], vp = VP[w]; uint64_t d0 = x & vp;
				d0 += ad; ad = d0 < ad; d0 += vp; ad |= d0 < vp;
				d0 ^= vp; d0 |= x;
				X[w] = x; D0[w] = d0; HN[w] = vp & d0; HP[w] = ~(vp | d0) | VN[w];
			}
			for (w = nword - 1, ad = 0; w >= 0; w--) {
				const uint64_t x = (D0[w] >> 1) | ad; ad = D0[w] << 63;
				X[w] = x; VN[w] = x & HP[w]; VP[w] = ~(x | HP[w]) | HN[w];
			}
		}
		if (!(D0[0] & 1ULL)) { ++err; if (err > cut) return; }
		if (i < tn0) {
			if (mode == 1 || mode == 2) { // running best end point of an extension (Levenshtein_distance.h:2758-2779 / 2889-2910)
				poff = i - thre; k = i + thre - pe;
				if (k >= 0) {
					if (tmp_e == INT32_MAX) { tmp_e = err; for (k = 0; poff < pe; poff++, k++) { tmp_e += hb_mw_bit(VP, k); tmp_e -= hb_mw_bit(VN, k); } }
					else { k = (thre << 1) - k; if (k >= 0) { tmp_e += hb_mw_bit(HP, k); tmp_e -= hb_mw_bit(HN, k); } }
					if (tmp_e <= ez.thre && tmp_e < ez.err) { ez.err = tmp_e; if (mode == 1) { ez.pe = pe; ez.te = i; } else { ez.ps = pidx - pe; ez.ts = tidx - i; } }
				}
			}
			for (k = 0; k < 4; k++) { // ed_infi_post_Peq
				uint64_t *pk = Peq + k * VS;
				for (int32_t w = 0; w + 1 < nword; w++) pk[w] = (pk[w] >> 1) | (pk[w + 1] << 63);
				pk[nword - 1] >>= 1;
			}
			++i_bd; c = 4;
			if (i_bd < pn) c = pch(i_bd);
			if (c < 4) Peq[c * VS + Peq_i] |= Peq_m;
		}
		uint64_t *o = ez.path + ez.pn;
		for (k = 0; k < nword; k++) { o[k] = D0[k]; o[nword + k] = VP[k]; o[2 * nword + k] = VN[k]; o[3 * nword + k] = HP[k]; o[4 * nword + k] = HN[k]; }
		ez.pn += (uint64_t)nword * 5;
	}
	if (mode == 0) {
		int32_t site = tn - 1 - thre; const int32_t ct = pn - 1;
		for (i = 0; site < ct; site++, i++) { err += hb_mw_bit(VP, i); err -= hb_mw_bit(VN, i); }
		if (site == ct && err <= thre) { ez.err = err; ez.pe = pn - 1; ez.te = tn - 1; }
		hb_mw_gen_trace(ez, thre, 1);
	} else if (mode == 1 || mode == 2) {
		int32_t site = tn - 1 - thre; const int32_t ct = pn - 1;
		for (i = 0; site < ct; i++) {
			err += hb_mw_bit(VP, i); err -= hb_mw_bit(VN, i); site++;
			if (err <= thre && err < ez.err) { ez.err = err; if (mode == 1) { ez.pe = site; ez.te = tn - 1; } else { ez.ps = pidx - site; ez.ts = tidx + 1 - tn; } }
		}
		if (err <= thre && err < ez.err) { ez.err = err; if (mode == 1) { ez.pe = site; ez.te = tn - 1; } else { ez.ps = pidx - site; ez.ts = tidx + 1 - tn; } }
		if (ez.te - ez.ts + 1 != tn) { ez.pn /= (uint64_t)tn; ez.pn *= (uint64_t)(ez.te + 1 - ez.ts); }
		if (mode == 1) hb_mw_gen_trace(ez, thre, 1);
		else {
			poff = ez.ps; ez.ps = pidx - ez.pe; ez.pe = pidx - poff;
			hb_mw_gen_trace(ez, thre, 0);
			poff = ez.ps; ez.ps = pidx - ez.pe; ez.pe = pidx - poff;
		}
	} else {
		int32_t site = tn - 1 - abs_diag, uge = INT32_MAX; const int32_t ai = pn - tn + abs_diag;
		for (i = 0; site < 0 && i < ai; i++, site++) { err += hb_mw_bit(VP, i); err -= hb_mw_bit(VN, i); }
		if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site; }
		site -= i;
		while (i < ai) {
			err += hb_mw_bit(VP, i); err -= hb_mw_bit(VN, i); ++i;
			if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site + i; }
			if (i == thre) uge = err;
		}
		if (uge <= thre && uge == ez.err) ez.pe = site + thre;
		hb_mw_gen_trace(ez, abs_diag, 1);
	}
}

template <int WPL> HB_HD_NI void hb_mw_align_w(int mode, const RdView &T, int64_t ps0, int32_t pn, const RdView &Q, int64_t qs0, int32_t tn, int32_t thre, int32_t abs_diag, MwEz &ez); // hb_mwalign_w.cuh

// lchain_refine, Hash_Table.cpp:2457-2541 with des = a (in place); t / p / f = DP scratch of a_n entries
HB_HD int64_t hb_lchain_refine(hb_hit_t *a, int64_t a_n, int64_t *t, int64_t *p, int32_t *f, int64_t max_skip, int64_t max_iter, int64_t max_dis, int64_t long_gap)
{
	if (a_n <= 0) return 0;
	int64_t max_f, n_skip, st, max_j, sc, msc, msc_i, dq, dr, dd, i, j, cL = 0;
	for (i = 1, f[0] = 0, p[0] = -1, msc_i = a_n - 1; i < a_n; i++) {
		j = i - 1;
		dq = (int64_t)a[i].self_offset - (int64_t)a[j].self_offset; dr = (int64_t)a[i].offset - (int64_t)a[j].offset;
		dd = dr > dq ? dr - dq : dq - dr;
		if (dd <= long_gap || dq > max_dis) { p[i] = i - 1; f[i] = (int32_t)i; }
		else break;
	}
	if (i < a_n) {
		for (j = 0; j < a_n; j++) t[j] = 0;
		f[0] = 0; p[0] = -1;
		for (i = 1, st = 0; i < a_n; ++i) {
			max_f = INT32_MIN; n_skip = 0; max_j = -1;
			if (i - st > max_iter) st = i - max_iter;
			j = i - 1;
			dq = (int64_t)a[i].self_offset - (int64_t)a[j].self_offset; dr = (int64_t)a[i].offset - (int64_t)a[j].offset;
			dd = dr > dq ? dr - dq : dq - dr;
			if (dd <= long_gap) dd = 0;
			sc = f[j] - dd;
			if (sc > max_f) { max_f = sc; max_j = j; }
			if (p[j] >= 0) t[p[j]] = i;
			for (--j; j >= st && (int64_t)a[i].self_offset <= max_dis + (int64_t)a[j].self_offset; --j) {
				dq = (int64_t)a[i].self_offset - (int64_t)a[j].self_offset; dr = (int64_t)a[i].offset - (int64_t)a[j].offset;
				dd = dr > dq ? dr - dq : dq - dr;
				if (dd <= long_gap) dd = 0;
				sc = f[j] - dd;
				if (sc > max_f) { max_f = sc; max_j = j; if (n_skip > 0) --n_skip; }
				else if (t[j] == (int32_t)i) { if (++n_skip > max_skip) break; }
				if (p[j] >= 0) t[p[j]] = i;
			}
			f[i] = (int32_t)max_f; p[i] = max_j;
		}
		i = a_n - 1; msc = f[i]; msc_i = i;
		for (j = i - 1; j >= 0 && (int64_t)a[i].self_offset <= max_dis + (int64_t)a[j].self_offset; --j)
			if (msc < f[j] && p[j] >= 0) { msc = f[j]; msc_i = j; }
	}
	i = msc_i; cL = 0;
	while (i >= 0) { t[cL++] = i; i = p[i]; }
	for (i = 0, j = cL - 1; i < j; i++, j--) { const int64_t x = t[i]; t[i] = t[j]; t[j] = x; }
	for (i = 0; i < cL; i++) a[i] = a[t[i]]; // t ascending and t[i] >= i: never reads an overwritten slot
	return cL;
}

// ---- window list of the overlap under construction ------------------------------------------------------------
HB_HD void hb_wc_push_trace(EcBCtx &C, uint32_t c, uint32_t len)
{
	c <<= 14;
	while (len >= 0x3fff) { if (C.wcn < C.wccap) C.wc[C.wcn] = (uint16_t)(c + 0x3fff); else C.ez.ovf = 1; C.wcn++; len -= 0x3fff; }
	if (len) { if (C.wcn < C.wccap) C.wc[C.wcn] = (uint16_t)(c + len); else C.ez.ovf = 1; C.wcn++; }
}
// ---- step C: reassign_gaps (Correct.cpp:25409-25430, row a11): move_wins (25274) left-normalises the indels of a window's
// cigar (adjust_gap 25167 slides every gap base left through matching bases, turning mismatches that become matches into
// matches) and turns leading / trailing mismatch runs into insertions (ajust_end_cigar 25252).  Runs on the finished
// cigar of the window that is being closed: in = C.wc, out = C.ez.cig, scratch = C.ez.path (both free at that point).
struct GapOut { uint16_t *c; int32_t n, cap; int ovf; };
HB_HD void hb_go_append(GapOut &o, uint32_t op, uint32_t l)
{ // append_cigar, Correct.cpp:25151-25165
	if (l == 0) return;
	if (o.n > 0 && (uint32_t)(o.c[o.n - 1] >> 14) == op) { l += o.c[o.n - 1] & 0x3fff; o.n--; }
	const uint32_t c = op << 14;
	while (l >= 0x3fff) { if (o.n < o.cap) o.c[o.n] = (uint16_t)(c + 0x3fff); else o.ovf = 1; o.n++; l -= 0x3fff; }
	if (l) { if (o.n < o.cap) o.c[o.n] = (uint16_t)(c + l); else o.ovf = 1; o.n++; }
}
HB_HD int hb_adjust_gap(EcBCtx &C, GapOut &o, int64_t pi, int64_t ti, uint32_t op0, uint16_t *buf, int64_t bcap, int64_t *rd_err)
{ // adjust_gap, Correct.cpp:25167-25250 (pstr = target on the overlap's strand, tstr = query)
	*rd_err = 0;
	if (o.n == 0) { hb_go_append(o, op0, 1); return 0; }
	if (op0 != 2 && op0 != 3) return 0;
	if (op0 == 2) ti--; else pi--;
	int64_t ci = o.n, op, cl, k, l0 = 0, l1 = 0, bn = 0; int ff = 0;
#define HB_BPUSH(v) do { if (bn < bcap) buf[bn] = (uint16_t)(v); else o.ovf = 1; bn++; } while (0)
	for (ci--; ci >= 0; ci--) {
		op = o.c[ci] >> 14; cl = o.c[ci] & 0x3fff;
		if (op == 2 || op == 3) { HB_BPUSH((op0 << 14) + 1); l0 = cl; HB_BPUSH((op << 14) + l0); break; }
		else if (op == 0) {
			for (k = cl - 1; k >= 0; k--, pi--, ti--) if (C.t.at(pi) != C.q.at(ti)) break;
			l1 = cl - k - 1; l0 = k + 1;
			if (l1 > 0) { HB_BPUSH((op << 14) + l1); ff = 1; }
			if (l0 > 0) { HB_BPUSH((op0 << 14) + 1); HB_BPUSH((op << 14) + l0); }
		} else {
			for (k = cl - 1, l0 = cl, l1 = 0; k >= 0; k--, pi--, ti--) {
				if (C.t.at(pi) == C.q.at(ti)) {
					l1 = l0 - k - 1; l0 = k;
					if (l1 > 0) { HB_BPUSH((op << 14) + l1); ff = 1; }
					HB_BPUSH(1); (*rd_err)++;
				}
			}
			if (l0 > 0) HB_BPUSH((op << 14) + l0);
			l0 = 0;
		}
		if (l0 > 0) break;
	}
#undef HB_BPUSH
	if (o.ovf) return 0;
	if (!ff) { hb_go_append(o, op0, 1); return 0; }
	if (ci >= 0) o.n = (int32_t)ci;
	else { o.n = 0; hb_go_append(o, op0, 1); }
	for (k = bn - 1; k >= 0; k--) hb_go_append(o, buf[k] >> 14, buf[k] & 0x3fff);
	return 1;
}
HB_HD int hb_adjust_end_cigar(hb_wl_t *p, uint16_t *ca, int32_t cn)
{ // ajust_end_cigar, Correct.cpp:25252-25272
	int rr = 0; int32_t ci;
	for (ci = 0; ci < cn; ci++) { const uint32_t op = ca[ci] >> 14, cl = ca[ci] & 0x3fff; if (op != 1) break; ca[ci] = (uint16_t)((3u << 14) + cl); p->y_start += (int32_t)cl; rr = 1; }
	for (ci = cn - 1; ci >= 0; ci--) { const uint32_t op = ca[ci] >> 14, cl = ca[ci] & 0x3fff; if (op != 1) break; ca[ci] = (uint16_t)((3u << 14) + cl); p->y_end -= (int32_t)cl; rr = 1; }
	return rr;
}
// move_wins for the aligned window p whose cigar is C.wc[0..C.wcn): returns the cigar to flush (C.wc itself when unchanged in place)
HB_HD const uint16_t *hb_move_wins(EcBCtx &C, hb_wl_t *p, int32_t *n_out)
{
	const uint16_t *cg = C.wc; const int32_t cn = C.wcn; int32_t ci; int mm;
	*n_out = cn;
	if (p->error == 0) return cg;
	for (ci = 0, mm = 0; ci < cn; ci++) { const uint32_t op = cg[ci] >> 14; if (op < 2) mm = 1; else if (mm) break; }
	if (ci >= cn) { hb_adjust_end_cigar(p, C.wc, cn); return cg; }
	GapOut o; o.c = C.gout ? C.gout : C.ez.cig; o.n = 0; o.cap = C.gout ? C.gcap : C.ez.ccap; o.ovf = 0;
	uint16_t *buf = (uint16_t *)C.ez.path; const int64_t bcap = (int64_t)(C.ez.pcap * 4);
	int64_t pi = p->y_start, ti = p->x_start, re;
	for (ci = 0, mm = 0; ci < cn && !o.ovf; ci++) {
		const uint32_t op = cg[ci] >> 14, cl = cg[ci] & 0x3fff;
		if (op < 2) { hb_go_append(o, op, cl); pi += cl; ti += cl; mm = 1; }
		else if (mm == 0) { hb_go_append(o, op, cl); if (op == 2) pi += cl; else ti += cl; }
		else for (uint32_t k = 0; k < cl && !o.ovf; k++) {
			if (hb_adjust_gap(C, o, pi, ti, op, buf, bcap, &re)) { p->error = (int16_t)(p->error - re); C.gap_re += re; }
			if (op == 2) pi++; else ti++;
		}
	}
	if (o.ovf) { C.ez.ovf = 1; return cg; }
	hb_adjust_end_cigar(p, o.c, o.n);
	*n_out = o.n;
	return o.c;
}

HB_HD void hb_b_flush(EcBCtx &C)
{ // the open window's cigar leaves the per-thread buffer for the shared pool (one contiguous piece, like aux_o->w_list.c)
	if (C.open < 0) return;
	hb_wl_t *p = &C.aw[C.open]; int32_t n = C.wcn; const uint16_t *src = C.wc;
	if (C.do_gaps && C.re_A != 0 && !C.ez.ovf) src = hb_move_wins(C, p, &n);
	if (!C.ez.ovf) {
#ifdef __CUDA_ARCH__
		const unsigned long long o = atomicAdd(C.pool_used, (unsigned long long)n);
#else
		const unsigned long long o = *C.pool_used; *C.pool_used += (unsigned long long)n;
#endif
		p->cidx = (uint32_t)o; p->clen = (uint32_t)n;
		if (o + (unsigned long long)n <= C.pool_cap) for (int32_t k = 0; k < n; k++) C.pool[o + k] = src[k];
	}
	C.open = -1; C.wcn = 0;
}
HB_HD void hb_push_alnw(EcBCtx &C, const AlnRes &ez)
{ // push_alnw + append_wcigar, Correct.cpp:15988-16019, 15954-15986
	hb_wl_t *p;
	if (C.awn > 0 && C.open == C.awn - 1 && C.wcn > 0) {
		p = &C.aw[C.awn - 1];
		const int64_t t = (int64_t)p->error + (int64_t)ez.err;
		if (p->x_end + 1 == ez.ts && p->y_end + 1 == ez.ps && t < INT16_MAX) {
			p->x_end = ez.te; p->y_end = ez.pe; p->error = (int16_t)(p->error + ez.err);
			if (ez.cn > 0 && !C.ez.ovf) {
				const uint32_t c0 = C.wc[C.wcn - 1] >> 14, c = ez.cig[0] >> 14; uint32_t l0 = C.wc[C.wcn - 1] & 0x3fff, l = ez.cig[0] & 0x3fff; int32_t ci = 1;
				for (; ci < ez.cn && (uint32_t)(ez.cig[ci] >> 14) == c; ci++) l += ez.cig[ci] & 0x3fff;
				if (c0 == c) { l += l0; C.wcn--; }
				hb_wc_push_trace(C, c, l);
				for (int32_t k = ci; k < ez.cn; k++) { if (C.wcn < C.wccap) C.wc[C.wcn] = ez.cig[k]; else C.ez.ovf = 1; C.wcn++; }
			}
			return;
		}
	}
	hb_b_flush(C);
	if (C.awn >= C.awcap) { C.ez.ovf = 1; return; }
	p = &C.aw[C.awn]; C.open = C.awn++;
	p->x_start = ez.ts; p->x_end = ez.te; p->y_start = ez.ps; p->y_end = ez.pe; p->extra_begin = p->extra_end = 0;
	p->error_threshold = 0; p->error = (int16_t)ez.err; p->cidx = 0; p->clen = (uint32_t)ez.cn;
	for (int32_t k = 0; k < ez.cn; k++) { if (C.wcn < C.wccap) C.wc[C.wcn] = ez.cig[k]; else C.ez.ovf = 1; C.wcn++; }
}
HB_HD AlnRes hb_aln_of(const MwEz &ez) { AlnRes r; r.ts = ez.ts; r.te = ez.te; r.ps = ez.ps; r.pe = ez.pe; r.err = ez.err; r.cn = ez.cn; r.cig = ez.cig; return r; }
HB_HD void hb_push_unmap_alnw(EcBCtx &C, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode)
{ // push_unmap_alnw, Correct.cpp:16021-16030
	hb_b_flush(C);
	if (C.awn >= C.awcap) { C.ez.ovf = 1; return; }
	hb_wl_t *p = &C.aw[C.awn++];
	p->x_start = (int32_t)qs; p->x_end = (int32_t)qe; p->y_start = (int32_t)ts; p->y_end = (int32_t)te;
	p->error_threshold = (int16_t)mode; p->error = INT16_MAX; p->extra_begin = p->extra_end = -1; p->cidx = p->clen = 0;
}
HB_HD void hb_set_exact(MwEz &ez, int64_t qs, int64_t qe, int64_t ts, int64_t te)
{ // set_exact_exz, Correct.cpp:16167-16175
	ez.thre = 0; ez.cn = 0; ez.err = 0; hb_mez_push_trace(ez, 0, (uint32_t)(qe - qs));
	ez.pl = (int32_t)(te - ts); ez.ps = (int32_t)ts; ez.pe = (int32_t)(ts + ez.pl - 1);
	ez.tl = (int32_t)(qe - qs); ez.ts = (int32_t)qs; ez.te = (int32_t)(qs + ez.tl - 1);
}
HB_HD int64_t hb_scale_ed_thre(uint32_t err, uint32_t max_err)
{ // scale_ed_thre, Correct.cpp:14354-14360
	uint64_t bd = ((uint64_t)err << 1) + 1, w = (bd >> 6) << 6; if (w < bd) w += 64;
	err = (uint32_t)((w - 1) >> 1); if (err > max_err) err = max_err;
	return err;
}
HB_HD void hb_adjust_ext_offset(int64_t *qs, int64_t *qe, int64_t *ts, int64_t *te, int64_t ql, int64_t tl, int64_t thre, int64_t mode)
{ // adjust_ext_offset, Correct.cpp:14400-14422
	int64_t qoff, toff;
	if (mode == 1) { qoff = ql - *qs; toff = tl - *ts; if (qoff <= toff) { *qe = ql; *te = *ts + qoff + thre; } else { *te = tl; *qe = *qs + toff + thre; } }
	else if (mode == 2) { qoff = *qe; toff = *te; if (qoff <= toff) { *qs = 0; *ts = *te - qoff - thre; } else { *ts = 0; *qs = *qe - toff - thre; } }
	if (*qs < 0) *qs = 0;
	if (*ts < 0) *ts = 0;
	if (*qe > ql) *qe = ql;
	if (*te > tl) *te = tl;
}

HB_HD int64_t hb_cal_estimate_err_hc(const EcZ &z, int64_t wl, int64_t qs, int64_t qe, int64_t ts, int64_t te, double e_rate, int64_t *exact)
{ // cal_estimate_err_hc, Correct.cpp:15403-15455 (z.w / z.wn = step A's window list)
	int64_t k, ws, we, wid, os, oe, ovlp, tot, cov_l, exa = 1, ots, ote, q0, t0; const int64_t est = (int64_t)((double)(qe - qs) * e_rate), wn = z.wn;
	*exact = 0;
	if (!wn) return est;
	if (qs < z.x_pos_s) qs = z.x_pos_s;
	if (qe > z.x_pos_e + 1) qe = z.x_pos_e + 1;
	ws = qs / wl; ws *= wl; wid = (ws - (z.x_pos_s / wl) * wl) / wl;
	if (wid >= wn) wid = wn - 1;
	for (k = wid; k < wn && qs > z.w[k].x_end; k++);
	if (k == wn) return est;
	for (; k >= 0 && qs < z.w[k].x_start; k--);
	if (k < 0) k = 0;
	for (tot = cov_l = 0, ots = ote = -1; k < wn && z.w[k].x_start < qe; k++) {
		if (z.w[k].y_end == -1) continue;
		ws = z.w[k].x_start; we = (int64_t)z.w[k].x_end + 1;
		os = qs > ws ? qs : ws; oe = qe < we ? qe : we;
		ovlp = oe > os ? oe - os : 0;
		if (!ovlp) continue;
		cov_l += ovlp;
		if (ovlp == we - ws) tot += z.w[k].error;
		else tot = (int64_t)((double)tot + ((double)z.w[k].error) * ((double)ovlp) / ((double)(we - ws)));
		if (z.w[k].error > 0) exa = 0;
		if (exa) {
			q0 = os - ws;
			we = (int64_t)z.w[k].y_end + 1; ws = we - ((int64_t)z.w[k].x_end + 1 - z.w[k].x_start);
			os = ts > ws ? ts : ws; oe = te < we ? te : we;
			ovlp = oe > os ? oe - os : 0;
			t0 = os - ws;
			if (ovlp && q0 == t0) {
				if (ote == -1) { ots = os; ote = oe; }
				else if (ote == os) ote = oe;
				else exa = 0;
			} else exa = 0;
		}
	}
	tot = (int64_t)((double)tot + (double)((qe - qs) - cov_l) * e_rate);
	if (exa && (qe - qs) == cov_l) { if ((qe - qs) == (ote - ots) && ots == ts && ote == te) *exact = 1; }
	return tot;
}

HB_HD uint32_t hb_fwd16(const uint8_t *p, uint64_t pos)
{ // 16 bases starting at `pos` as one word of 2-bit codes, first base in the top two bits (reads are padded: the 5-byte window never leaves the store)
	const uint8_t *b = p + (pos >> 2); const uint32_t sh = (uint32_t)(pos & 3) << 1;
	const uint64_t v = (uint64_t)b[0] << 32 | (uint64_t)b[1] << 24 | (uint64_t)b[2] << 16 | (uint64_t)b[3] << 8 | b[4];
	return (uint32_t)(v >> (8 - sh));
}
HB_HD uint32_t hb_rc16(uint32_t w)
{ // reverse the 16 2-bit groups and complement
#ifdef __CUDA_ARCH__
	w = __brev(w);
#else
	w = (w >> 16) | (w << 16); w = ((w & 0xff00ff00u) >> 8) | ((w & 0x00ff00ffu) << 8); w = ((w & 0xf0f0f0f0u) >> 4) | ((w & 0x0f0f0f0fu) << 4);
	w = ((w & 0xccccccccu) >> 2) | ((w & 0x33333333u) << 2); w = ((w & 0xaaaaaaaau) >> 1) | ((w & 0x55555555u) << 1);
#endif
	w = ((w >> 1) & 0x55555555u) | ((w & 0x55555555u) << 1);
	return ~w;
}
// query[qs, qs+n) == target-on-strand[ts, ts+n) ?  16 bases per step when neither read holds an N
HB_HD bool hb_seq_equal(const RdView &Q, int64_t qs, const RdView &T, int64_t ts, int64_t n)
{
	int64_t k = 0;
	if (!Q.nn && !T.nn) {
		for (; k + 16 <= n; k += 16) {
			const uint32_t a = hb_fwd16(Q.p, (uint64_t)(qs + k));
			const uint32_t b = T.rev ? hb_rc16(hb_fwd16(T.p, (uint64_t)((int64_t)T.len - (ts + k) - 16))) : hb_fwd16(T.p, (uint64_t)(ts + k));
			if (a != b) return false;
		}
	}
	for (; k < n; k++) if (Q.at(qs + k) != T.at(ts + k)) return false;
	return true;
}
HB_HD int64_t hb_cal_exact(EcBCtx &C, const EcZ &z, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode)
{ // cal_exact_exz, Correct.cpp:15725-15762 (memcmp on decoded strings: an N only equals an N)
	MwEz &ez = C.ez; int64_t ql = qe - qs, tl;
	ez.err = INT32_MAX; ez.thre = 0; ez.cn = 0;
	if (mode == 3) { ts = (qs - z.x_pos_s) + z.y_pos_s; ts += hb_y_start_offset(qs, z.fc, z.fc_n, &C.bad); te = ts + ql; }
	else if (mode == 1) te = ts + ql;
	else if (mode == 2) ts = te - ql;
	if (ts < 0) ts = 0;
	if (ts > C.tl) ts = C.tl;
	if (te > C.tl) te = C.tl;
	ql = qe - qs; tl = te - ts;
	if (ql != tl) return 0;
	if (!hb_seq_equal(C.q, qs, C.t, ts, ql)) return 0;
	ez.err = 0; hb_mez_push_trace(ez, 0, (uint32_t)ql);
	ez.pl = (int32_t)tl; ez.ps = (int32_t)ts; ez.pe = (int32_t)(ts + tl - 1);
	ez.tl = (int32_t)ql; ez.ts = (int32_t)qs; ez.te = (int32_t)(qs + ql - 1);
	return 1;
}

HB_HD int64_t hb_cal_exz_adv(EcBCtx &C, const EcZ &z, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t thre, int64_t *pthre, int64_t mode)
{ // cal_exz_infi_adv, Correct.cpp:15617-15666
	MwEz &ez = C.ez; int64_t aux_beg = 0, ql = qe - qs, tl = te - ts, dd = ql > tl ? ql : tl;
	ez.err = INT32_MAX;
	if (mode == 3) { // update_semi_coord, Correct.cpp:14364-14381
		const int64_t th = thre > dd ? dd : thre, aln_l = (qe - qs) + (th << 1); int64_t aux_end, l;
		ts = (qs - z.x_pos_s) + z.y_pos_s; ts += hb_y_start_offset(qs, z.fc, z.fc_n, &C.bad);
		if (!hb_init_waln(th, ts, C.tl, aln_l, &aux_beg, &aux_end, &ts, &l)) ts = te = aux_beg = -1;
		else te = ts + l;
	} else if (mode == 1 || mode == 2) hb_adjust_ext_offset(&qs, &qe, &ts, &te, C.ql, C.tl, thre > dd ? dd : thre, mode);
	if (qe > qs && te > ts && ts != -1 && te != -1) {
		ql = qe - qs; tl = te - ts; dd = ql > tl ? ql : tl;
		if (thre > dd) thre = dd;
		if (thre <= *pthre) return 0;
		*pthre = thre;
		if (ez.warp) { // the band over the lanes: one word per lane up to 32 words, two beyond
			if ((((int32_t)thre << 1) + 1 + 63) >> 6 <= 32) hb_mw_align_w<1>((int)mode, C.t, ts, (int32_t)tl, C.q, qs, (int32_t)ql, (int32_t)thre, (int32_t)aux_beg, ez);
			else hb_mw_align_w<2>((int)mode, C.t, ts, (int32_t)tl, C.q, qs, (int32_t)ql, (int32_t)thre, (int32_t)aux_beg, ez);
		} else hb_mw_align((int)mode, C.t, ts, (int32_t)tl, C.q, qs, (int32_t)ql, (int32_t)thre, (int32_t)aux_beg, ez);
		if (ez.ovf) return 0;
		if (ez.err <= ez.thre) { ez.ps += (int32_t)ts; ez.pe += (int32_t)ts; ez.ts += (int32_t)qs; ez.te += (int32_t)qs; return 1; }
	}
	return 0;
}

// hc_aln_exz_adv_hc, Correct.cpp:16178-16260 (maxl = MAX_SIN_L, maxe = MAX_SIN_E, force_l = FORCE_SIN_L, estimate_err = -1) without its
// push_alnw: returns 1 = aligned (C.ez holds the result), 2 = empty segment (nothing to push), 0 = not aligned (or C.ez.ovf)
// CONV (device, thread-per-segment kernels): every lane of the warp calls this together (`live` = the lane has a segment) and nobody leaves before the
// end; the lanes meet (hb_wsync) in front of every aligner call, so that the aligner — a function call the compiler does not reconverge the warp for: ncu
// showed it entered 2.8 times per warp with 11 lanes each — runs once per warp and step with every lane that needs it.
// NOALN (the pre-pass kernel): stops where an alignment would start (returns 5), with no aligner code in the instantiation at all.
template <bool CONV, bool NOALN = false>
HB_HD int hb_seg_align_t(EcBCtx &C, const EcZ &z, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode, bool live)
{
	MwEz &ez = C.ez; int64_t thre = 0, thre0 = -1, pthre = -1, full = 0, est = 0; const int64_t ql = qe - qs;
	int ret = -1; // undecided
	ez.err = INT32_MAX; ez.thre = 0;
	if (!live) ret = 0;
	if (ret < 0) {
		if (ts == -1 && te == -1) mode = 3;
		if (ql == 0 && te - ts == 0) ret = 2;
		else if (ql <= 0 || te - ts <= 0) ret = 0;
	}
	if (ret < 0) {
		est = hb_cal_estimate_err_hc(z, C.w_l, qs, qe, ts, te, C.e_rate, &full);
		if (est == 0) {
			if (full) { hb_set_exact(ez, qs, qe, ts, te); ret = 1; }
			else if (hb_cal_exact(C, z, qs, qe, ts, te, mode)) ret = 1;
		}
	}
	bool want = ret < 0 && ql <= HB_MAX_SIN_L && (est >> 1) <= HB_MAX_SIN_E;
	if (NOALN) return want ? 5 : (ret < 0 ? 0 : ret);
	if (want && C.no_myers) { ret = 5; want = false; }
	if (!CONV && !want) return ret < 0 ? 0 : ret;
	// thresholds in the reference's order: the estimate, len*e_rate, twice that, 0.51*len, and the maximum for short segments; each one
	// (but the first and the last) only if it exceeds the one before — one call site, so the aligner exists once in the kernel
	for (int step = 0; step < 5; step++) {
		bool go = false;
		if (want) {
			go = true;
			if (step == 0) thre = hb_scale_ed_thre((uint32_t)est, HB_MAX_SIN_E);
			else if (step == 1) { thre0 = thre; thre = (int64_t)((double)ql * C.e_rate); thre = hb_scale_ed_thre((uint32_t)thre, HB_MAX_SIN_E); go = thre > thre0; }
			else if (step == 2) { thre0 = thre; thre <<= 1; thre = hb_scale_ed_thre((uint32_t)thre, HB_MAX_SIN_E); go = thre > thre0; }
			else if (step == 3) { thre0 = thre; thre = (int64_t)((double)ql * 0.51); thre = hb_scale_ed_thre((uint32_t)thre, HB_MAX_SIN_E); go = thre > thre0; }
			else { go = ql <= HB_FORCE_SIN_L; thre = HB_MAX_SIN_E; }
		}
		if (CONV) hb_wsync();
		if (go) {
			if (hb_cal_exz_adv(C, z, qs, qe, ts, te, thre, &pthre, mode)) { ret = 1; want = false; }
			else if (ez.ovf) { ret = 0; want = false; }
		}
	}
	return ret < 0 ? 0 : ret;
}
HB_HD int hb_seg_align(EcBCtx &C, const EcZ &z, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode) { return hb_seg_align_t<false>(C, z, qs, qe, ts, te, mode, true); }

// ---- the three pieces of step B ---------------------------------------------------------------------------------
// prep   : return_t_chain (refine the chain in place) + hc_ovlp_base_direct's whole-overlap exact shortcut (Correct.cpp:17430-17459)
// segment: one inter-anchor segment i of [0, ch_n]: coordinates (17470-17490) + hb_seg_align         -> EcSeg (independent of its neighbours)
// merge  : push_alnw / push_unmap_alnw over the segments in order, reassign_gaps, totals, update_overlap_region
struct EcPrep { int32_t ch_n, shortcut; int32_t q0, q1, t0, t1; }; // shortcut: one exact window [q0,q1) x [t0,t1)
// status: 0 unmapped (ts,te,ps,pe = q0,q1-1,t0,t1-1), 1 aligned with its cigar at coff (cn runs) in the segment pool, 2 empty,
//         3 aligned without error (cigar = one match run, not stored), 4 deferred (needs more scratch)
struct EcSeg { int32_t ts, te, ps, pe, err; uint32_t coff; uint16_t cn; uint8_t status, mode; };

HB_HD void hb_ecb_prep(const EcZ &zA, int64_t re_A, int64_t ql, int64_t tl, hb_hit_t *ch_a, int64_t scn, int64_t *dp_t, int64_t *dp_p, int32_t *dp_f, EcPrep *pr)
{
	const int64_t ch_n = hb_lchain_refine(ch_a, scn, dp_t, dp_p, dp_f, 50, 5000, 512, 16);
	for (int64_t i = ch_n; i < scn; i++) ch_a[i].id_strand = (ch_a[i].id_strand & 0x80000000u) | 0x7fffffffu;
	pr->ch_n = (int32_t)ch_n; pr->shortcut = 0; pr->q0 = pr->q1 = pr->t0 = pr->t1 = 0;
	if (ch_n > 0 && re_A == 0 && zA.wn) {
		const int32_t zn = zA.wn; int32_t k; int64_t q[2], t[2];
		for (k = 1; k < zn; k++) {
			if (zA.w[k].error == 0 && zA.w[k - 1].error == 0 && zA.w[k].x_start == zA.w[k - 1].x_end + 1 && zA.w[k].y_end == zA.w[k - 1].y_end + (zA.w[k].x_end - zA.w[k - 1].x_end)) continue;
			break;
		}
		if (k >= zn) {
			q[0] = zA.w[0].x_start; q[1] = zA.w[zn - 1].x_end; t[1] = zA.w[zn - 1].y_end; t[0] = (int64_t)zA.w[0].y_end - (zA.w[0].x_end - zA.w[0].x_start);
			if (q[0] <= t[0]) { t[0] -= q[0]; q[0] = 0; } else { q[0] -= t[0]; t[0] = 0; }
			const int64_t qr = ql - q[1] - 1, tr = tl - t[1] - 1;
			if (qr <= tr) { q[1] = ql - 1; t[1] += qr; } else { t[1] = tl - 1; q[1] += tr; }
			if (q[0] == zA.w[0].x_start && q[1] == zA.w[zn - 1].x_end) { pr->shortcut = 1; pr->q0 = (int32_t)q[0]; pr->q1 = (int32_t)(q[1] + 1); pr->t0 = (int32_t)t[0]; pr->t1 = (int32_t)(t[1] + 1); }
		}
	}
}
// segment i of an overlap with refined chain ch_a[0..ch_n): returns hb_seg_align's status; unmapped coordinates in uq / ut
template <bool CONV, bool NOALN = false>
HB_HD int hb_ecb_segment_t(EcBCtx &C, const EcZ &zA, const hb_hit_t *ch_a, int64_t ch_n, int64_t i, int64_t *uq, int64_t *ut, int64_t *umode, bool live)
{
	int64_t q[2], t[2], mode = 3; const int64_t l = i - 1;
	q[0] = q[1] = t[0] = t[1] = -1;
	if (live) {
		if (l >= 0) { q[0] = ch_a[l].self_offset; t[0] = ch_a[l].offset; } else q[0] = 0;
		if (i < ch_n) { q[1] = ch_a[i].self_offset; t[1] = ch_a[i].offset; } else q[1] = C.ql;
		if (t[0] != -1 && t[1] != -1) mode = 0;
		else if (t[0] != -1 && t[1] == -1) mode = 1;
		else if (t[0] == -1 && t[1] != -1) mode = 2;
		else mode = 3;
		if (mode == 1 || mode == 2) hb_adjust_ext_offset(&q[0], &q[1], &t[0], &t[1], C.ql, C.tl, 0, mode);
	}
	uq[0] = q[0]; uq[1] = q[1]; ut[0] = t[0]; ut[1] = t[1]; *umode = mode;
	return hb_seg_align_t<CONV, NOALN>(C, zA, q[0], q[1], t[0], t[1], mode, live);
}
HB_HD int hb_ecb_segment(EcBCtx &C, const EcZ &zA, const hb_hit_t *ch_a, int64_t ch_n, int64_t i, int64_t *uq, int64_t *ut, int64_t *umode) { return hb_ecb_segment_t<false>(C, zA, ch_a, ch_n, i, uq, ut, umode, true); }
// what hc_ovlp_base_direct does with a segment's result
HB_HD void hb_ecb_apply(EcBCtx &C, int status, const AlnRes &r, const int64_t *uq, const int64_t *ut, int64_t umode)
{
	if (status == 1) hb_push_alnw(C, r);
	else if (status == 0) hb_push_unmap_alnw(C, uq[0], uq[1] - 1, ut[0], ut[1] - 1, umode);
}
// totals, rechain flag, update_overlap_region (Correct.cpp:17857-17866, 17676, 17249-17275) once every window is closed
HB_HD void hb_ecb_finish(EcBCtx &C, const EcZ &zA, int64_t re_A, hb_alnb_t *out)
{
	const int64_t ql = C.ql, tl = C.tl; int64_t i;
	hb_b_flush(C);
	out->need_rechain = 0; out->re = 0; out->w_n = 0; out->nh_err = re_A;
	if (C.ez.ovf) { out->st = -1; return; }
	int64_t tot_e = 0, xs = zA.x_pos_s, xe = zA.x_pos_e, ys = zA.y_pos_s, ye = 0;
	for (i = 0; i < C.awn; i++) {
		const hb_wl_t &u = C.aw[i];
		if (u.error == INT16_MAX && u.clen == 0 && u.extra_end < 0) {
			const int64_t xl = (int64_t)u.x_end + 1 - u.x_start, yl = (int64_t)u.y_end + 1 - u.y_start;
			if (xl >= HB_FORCE_SIN_L && yl >= HB_FORCE_SIN_L) out->need_rechain = 1;
			tot_e += xl >= yl ? xl : yl;
		} else tot_e += u.error;
	}
	if (C.awn) { xs = C.aw[0].x_start; xe = C.aw[C.awn - 1].x_end; ys = C.aw[0].y_start; ye = C.aw[C.awn - 1].y_end; }
	if (xs <= ys) { ys -= xs; xs = 0; } else { xs -= ys; ys = 0; }
	{ const int64_t xr = ql - xe - 1, yr = tl - ye - 1; if (xr <= yr) { xe = ql - 1; ye += xr; } else { ye = tl - 1; xe += yr; } }
	out->st = 2; out->re = tot_e + C.gap_re; out->w_n = (uint32_t)C.awn; out->nh_err = re_A - C.gap_re; // re = step B's total (before step C)
	out->x_pos_s = (uint32_t)xs; out->x_pos_e = (uint32_t)xe; out->y_pos_s = (uint32_t)ys; out->y_pos_e = (uint32_t)ye;
	if (C.bad) out->st = -2;
}
HB_HD void hb_ecb_begin(EcBCtx &C, int64_t re_A) { C.no_myers = 0; C.awn = 0; C.wcn = 0; C.open = -1; C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0; C.re_A = re_A; C.gap_re = 0; }

// One accepted overlap.  zA = the overlap with step A's window list; re_A = step A's error estimate; ch_a / ch_n = its chain
// anchors (refined in place, dropped anchors get id 0x7fffffff like return_t_chain does).  Result: out->st = 2 done,
// -1 deferred (scratch too small: nothing of this overlap is valid); out->need_rechain = 1 when an unaligned window of
// >= FORCE_SIN_L remains, which the reference re-seeds (rechain_aln_hc, Correct.cpp:17669 — not built yet).
HB_HD void hb_ec_overlap_B(EcBCtx &C, const EcZ &zA, int64_t re_A, hb_hit_t *ch_a, int64_t scn, int64_t *dp_t, int64_t *dp_p, int32_t *dp_f, hb_alnb_t *out)
{ // the three pieces run back to back by one thread (host emulation and the reference for the parallel pipeline's results)
	EcPrep pr; int64_t uq[2], ut[2], um;
	hb_ecb_begin(C, re_A);
	hb_ecb_prep(zA, re_A, C.ql, C.tl, ch_a, scn, dp_t, dp_p, dp_f, &pr);
	if (pr.shortcut) { hb_set_exact(C.ez, pr.q0, pr.q1, pr.t0, pr.t1); hb_push_alnw(C, hb_aln_of(C.ez)); }
	else for (int64_t i = 0; i <= pr.ch_n && pr.ch_n > 0 && !C.ez.ovf; i++) {
		const int st = hb_ecb_segment(C, zA, ch_a, pr.ch_n, i, uq, ut, &um);
		if (C.ez.ovf) break;
		hb_ecb_apply(C, st, hb_aln_of(C.ez), uq, ut, um);
	}
	hb_ecb_finish(C, zA, re_A, out);
}

// ---- storing / loading a segment's result between the segment and the merge kernels -------------------------------
HB_HD void hb_seg_store(EcBCtx &C, int st, const int64_t *uq, const int64_t *ut, int64_t um, EcSeg *sg, uint16_t *spool, unsigned long long *spool_used, uint64_t spool_cap)
{
	sg->mode = (uint8_t)um; sg->cn = 0; sg->coff = 0; sg->err = 0; sg->ts = sg->te = sg->ps = sg->pe = 0;
	if (C.ez.ovf) { sg->status = 4; return; }
	if (st == 0) { sg->status = 0; sg->ts = (int32_t)uq[0]; sg->te = (int32_t)(uq[1] - 1); sg->ps = (int32_t)ut[0]; sg->pe = (int32_t)(ut[1] - 1); return; }
	if (st == 2) { sg->status = 2; return; }
	const MwEz &ez = C.ez;
	sg->ts = ez.ts; sg->te = ez.te; sg->ps = ez.ps; sg->pe = ez.pe; sg->err = ez.err; sg->cn = (uint16_t)ez.cn;
	if (ez.cn == 1) { sg->status = 3; sg->coff = ez.cig[0]; return; } // a single run travels inside the record
	if (ez.cn > 0xffff) { sg->status = 4; return; }
#ifdef __CUDA_ARCH__
	const unsigned long long o = atomicAdd(spool_used, (unsigned long long)ez.cn);
#else
	const unsigned long long o = *spool_used; *spool_used += (unsigned long long)ez.cn;
#endif
	sg->status = 1; sg->coff = (uint32_t)o;
	if (o + (unsigned long long)ez.cn <= spool_cap && o + (unsigned long long)ez.cn < (1ull << 32)) for (int32_t k = 0; k < ez.cn; k++) spool[o + k] = ez.cig[k];
}
// merge: replay one overlap's stored segments (Correct.cpp:17470-17507 with the alignments already done)
HB_HD void hb_ecb_merge(EcBCtx &C, const EcZ &zA, int64_t re_A, const EcPrep &pr, const EcSeg *segs, const uint16_t *spool, hb_alnb_t *out)
{
	hb_ecb_begin(C, re_A);
	if (pr.shortcut) { hb_set_exact(C.ez, pr.q0, pr.q1, pr.t0, pr.t1); hb_push_alnw(C, hb_aln_of(C.ez)); }
	else for (int32_t i = 0; i <= pr.ch_n && pr.ch_n > 0 && !C.ez.ovf; i++) {
		const EcSeg sg = segs[i]; uint16_t one;
		if (sg.status == 1 || sg.status == 3) {
			AlnRes r; r.ts = sg.ts; r.te = sg.te; r.ps = sg.ps; r.pe = sg.pe; r.err = sg.err; r.cn = sg.cn;
			if (sg.status == 3) { one = (uint16_t)sg.coff; r.cig = &one; } else r.cig = spool + sg.coff;
			hb_push_alnw(C, r);
		} else if (sg.status == 0) hb_push_unmap_alnw(C, sg.ts, sg.te, sg.ps, sg.pe, sg.mode);
		else if (sg.status == 4) C.ez.ovf = 1; // a segment no tier could align: the overlap stays deferred (reported by the host)
	}
	hb_ecb_finish(C, zA, re_A, out);
}
#include "hb_mwalign_w.cuh"
